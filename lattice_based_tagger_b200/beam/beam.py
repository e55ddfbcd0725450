"""Result object of the beam search (host side).

`Sequence` is what `Tagger.tag` returns: the word list BOS .. EOS, the path score and the number
of trailing unknown words — the same three attributes as the reference's `Sequence`
(`beam/beam.py:88-124`).  The search itself (`beam_search`, `Beam`; `beam/beam.py:5-86`) is the
device kernel `csrc/beam.cuh`; hypotheses there are back-pointer entries, and only the best path
is materialised as a `Sequence`.
"""


class Sequence:
    def __init__(self, sequences, score, num_unk=0):
        self.sequences = sequences
        self.score = score
        self.num_unk = num_unk

    def __repr__(self):
        words = '[\n    {}\n  ]'.format('\n    '.join(str(w) for w in self.sequences))
        return 'Sequences(\n  words : {}\n  score : {}\n  num unks in tails : {}\n)'.format(
            words, self.score, self.num_unk)

    __str__ = __repr__
