"""`analyze_morphology` (host side): the reference's lemmatizer entry point
(`dictionary/lemmatizer.py:5-51`) answered by the lattice kernel.

The candidates `get_lemma_candidates` enumerates (`lemmatizer.py:53-112`) — plain splits, the rules
of the one-syllable key once per rule of that key, the rules of the two- and three-syllable keys in
set order — and the dictionary checks of `analyze_morphology` are what `csrc/lattice.cuh` does for
every substring it looks up; one word is a batch of one, looked up in `LT_LOOKUP_EXACT` mode.
"""

from ..tagset import Adjective, Eomi, Verb

_ENGINES = {}
_MAX_ENGINES = 4


def _lemma_lookup(verbs, adjectives, eomis, lemma_rules, device=0):
    from .dictionary import MorphemeDictionary
    from .lookup import ExactLookup
    key = (id(verbs), id(adjectives), id(eomis), id(lemma_rules), device)
    sizes = (len(verbs), len(adjectives), len(eomis), len(lemma_rules))
    hit = _ENGINES.get(key)
    if hit is not None and hit[0] == sizes and hit[1] is verbs and hit[2] is eomis:
        return hit[3]
    sets = {Verb: set(verbs), Adjective: set(adjectives), Eomi: set(eomis)}
    dictionary = MorphemeDictionary({tag: morphs for tag, morphs in sets.items() if morphs}, lemma_rules)
    lookup = ExactLookup(dictionary, device=device)
    while len(_ENGINES) >= _MAX_ENGINES:
        _ENGINES.pop(next(iter(_ENGINES)))[3]._release()
    _ENGINES[key] = (sizes, verbs, eomis, lookup)
    return lookup


def analyses_of(words):
    """Two-morpheme `Word`s -> [((stem, tag), (eomi, 'Eomi')), ...] in lemmatizer order."""
    return [((w.morph0, w.tag0), (w.morph1, w.tag1)) for w in words if w.tag1 is not None]


def analyze_morphology(word, verbs, adjectives, eomis, lemma_rules, debug=False):
    """
    >>> analyze_morphology('파랬다', {}, {'파랗'}, {'았다'}, {'랬': (('랗', '았'),)})
    [(('파랗', 'Adjective'), ('았다', 'Eomi'))]
    """
    if not word:
        return []
    return analyses_of(_lemma_lookup(verbs, adjectives, eomis, lemma_rules).lookup(word))
