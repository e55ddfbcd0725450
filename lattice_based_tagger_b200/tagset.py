"""Tag vocabulary of the lattice tagger.

Mirrors the 13 tag strings of the reference (`lattice_tagger/tagset.py:1-15`): ten
part-of-speech tags plus the BOS / EOS sentinels and the unknown-word tag.  On the device every
tag is a small integer id; ids 0..12 are fixed (`TAG_IDS`), further tag names that appear in a
user dictionary are numbered 13.. by the table compiler (`compile.py`).
"""

Noun = 'Noun'
Pronoun = 'Pronoun'
Number = 'Number'
Josa = 'Josa'
Adjective = 'Adjective'
Verb = 'Verb'
Eomi = 'Eomi'
Adverb = 'Adverb'
Determiner = 'Determiner'
Exclamation = 'Exclamation'

BOS = 'BOS'
EOS = 'EOS'

Unk = 'Unknown'

#: fixed device ids of the built-in tags (order of `lattice_tagger/tagset.py`)
BUILTIN_TAGS = (Noun, Pronoun, Number, Josa, Adjective, Verb, Eomi, Adverb, Determiner,
                Exclamation, BOS, EOS, Unk)
TAG_IDS = {tag: idx for idx, tag in enumerate(BUILTIN_TAGS)}

#: tags that take part in the contextual feature (template 8, `features/feature.py:88`)
CONTEXTUAL_TAGS = (Noun, Adverb, Adjective, Verb)

#: upper bound on distinct tags a compiled table can hold (29-bit tag sets next to 3 lemma bits)
MAX_TAGS = 29

__all__ = ['Noun', 'Pronoun', 'Number', 'Josa', 'Adjective', 'Verb', 'Eomi', 'Adverb',
           'Determiner', 'Exclamation', 'BOS', 'EOS', 'Unk']
