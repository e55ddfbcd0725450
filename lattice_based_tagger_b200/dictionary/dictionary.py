"""Dictionary containers and loaders (host side).

These mirror the reference's dictionary plug-in protocol — `Word` (`dictionary/dictionary.py:169`),
`WordDictionary` (`:202-262`), `MorphemeDictionary` (`:265-315`), `load_dictionary` (`:349-360`),
`load_rules` (`:365-378`) and the bundled `Demo*/Base*` dictionaries (`:318-347`) — so that user
code written against `lattice_tagger.dictionary` keeps working.  They are *containers*: the
string lookups, the lemmatizer and the eojeol lattice enumeration that the reference performs in
Python on these sets run on the GPU (`csrc/lattice.cuh`) from the tables `compile.py` builds
out of a container.  Nothing here scores or searches.

`Tagger` also accepts the reference's own dictionary objects: only the attributes
`tag_to_morphs`, `rules`, `verbs`, `adjectives`, `eomis` are read.
"""

import glob
import os
from collections import namedtuple

from ..tagset import Adjective, Eomi, Verb

_WORD_FIELDS = ('word', 'morph0', 'morph1', 'tag0', 'tag1', 'len', 'b', 'e', 'is_l')


class Word(namedtuple('Word', _WORD_FIELDS)):
    """One lattice edge: surface `word`, one or two (morph, tag) pairs, a free `len` field, the
    syllable span [b, e) in the space-stripped sentence and the left-of-eojeol flag.

    Same nine fields, order and printed form as the reference's `Word`
    (`dictionary/dictionary.py:169-199`).  `len` is not always `e - b`.
    """

    __slots__ = ()

    def __str__(self):
        left = ', L' if self.is_l else ''
        if self.morph1:
            analysis = f'{self.morph0}/{self.tag0} + {self.morph1}/{self.tag1}'
        else:
            analysis = f'{self.morph0}/{self.tag0}'
        return f'Word({self.word}, {analysis}, len={self.len:d}, b={self.b:d}, e={self.e:d}{left})'

    __repr__ = __str__


class WordDictionary:
    """tag -> set of surface forms (reference `WordDictionary`, `dictionary.py:202-262`).

    Iteration order of `tag_to_morphs` is observable: it is the order in which a string's tags
    are emitted into the lattice (`get_tags`), so it is kept as given.
    """

    def __init__(self, tag_to_morphs):
        self.tag_to_morphs = tag_to_morphs

    def check(self, morph, tag):
        return morph in self.tag_to_morphs.get(tag, ())

    def get_tags(self, morph):
        return [tag for tag, morphs in self.tag_to_morphs.items() if morph in morphs]

    def add(self, morphs, tag, force=False):
        if isinstance(morphs, str):
            morphs = {morphs}
        if tag not in self.tag_to_morphs:
            if not force:
                raise ValueError('{} tag does not exist in dictionary'.format(tag))
            self.tag_to_morphs[tag] = set()
        self.tag_to_morphs[tag].update(morphs)
        self._invalidate()

    def _invalidate(self):
        # device tables built for lookup() / lemmatize() describe the dictionary as it was
        lookup = self.__dict__.pop('_exact_lookup', None)
        if lookup is not None:
            lookup._release()

    def remove_words(self, morphs, tag):
        if isinstance(morphs, str):
            morphs = {morphs}
        if tag not in self.tag_to_morphs:
            raise ValueError('{} tag does not exist in dictionary'.format(tag))
        drop = set(morphs)
        # the reference rebinds the tag's set (dictionary.py:260-262); `verbs` / `adjectives` /
        # `eomis` of a MorphemeDictionary keep pointing at the old object, and the table
        # compiler reads them separately for exactly that reason.
        self.tag_to_morphs[tag] = {m for m in self.tag_to_morphs[tag] if m not in drop}
        self._invalidate()


class MorphemeDictionary(WordDictionary):
    """Morpheme dictionary with conjugation rules (reference `dictionary.py:265-315`).

    `rules` maps a conjugated surface of one to three syllables to a tuple of
    `(stem, eomi)` canonical forms; the tuple order is observable.
    """

    def __init__(self, tag_to_morph, rules=None):
        super().__init__(tag_to_morph)
        self.rules = {} if rules is None else rules
        self.verbs = tag_to_morph.get(Verb, {})
        self.adjectives = tag_to_morph.get(Adjective, {})
        self.eomis = tag_to_morph.get(Eomi, {})

    def _exact(self):
        lookup = self.__dict__.get('_exact_lookup')
        if lookup is None:
            from .lookup import ExactLookup
            lookup = self.__dict__['_exact_lookup'] = ExactLookup(self)
        return lookup

    def lookup(self, word, b=0, is_l=False):
        """One `Word` per tag that lists `word` (dictionary order), then its lemmatised analyses
        (reference `dictionary.py:304-312`) — looked up on the device (`LT_LOOKUP_EXACT`)."""
        if not word:
            return []
        return [w._replace(b=w.b + b, e=w.e + b, is_l=is_l) for w in self._exact().lookup(word)]

    def lemmatize(self, word):
        """[((stem, tag), (eomi, 'Eomi')), ...] (reference `dictionary.py:314-315`), from the device."""
        if not word:
            return []
        from .lemmatizer import analyses_of
        return analyses_of(self._exact().lookup(word))


def load_dictionary(directory):
    """`<directory>/<Tag>.txt` -> {tag: set(first column)} (reference `dictionary.py:349-360`).

    Tag order is `glob` order, as in the reference, because that order is what a string's tags
    are emitted in.
    """
    tag_to_morphs = {}
    for path in glob.glob('%s/*.txt' % directory):
        tag = path.split('/')[-1][:-4]
        with open(path, encoding='utf-8') as f:
            tag_to_morphs[tag] = {line.split()[0] for line in f}
    return tag_to_morphs


def load_rules(path):
    """Three-column rule file `surface stem eomi` -> {surface: tuple((stem, eomi), ...)}
    (reference `dictionary.py:365-378`).  Malformed lines are reported and skipped."""
    collected = {}
    with open(path, encoding='utf-8') as f:
        for lineno, line in enumerate(f):
            columns = line.split()
            if not columns:
                continue
            if len(columns) != 3:
                print('Exception (%d line) : %s' % (lineno, line))
                continue
            surface, stem, eomi = columns
            collected.setdefault(surface, set()).add((stem, eomi))
    return {surface: tuple(canons) for surface, canons in collected.items()}


def write_rules(rules, path):
    """Inverse of `load_rules` (reference `dictionary.py:380-384`)."""
    with open(path, 'w', encoding='utf-8') as f:
        for surface, canons in rules.items():
            for stem, eomi in canons:
                f.write('%s %s %s\n' % (surface, stem, eomi))


def find_resources(name):
    """Locate a bundled dictionary directory (`base`, `demo_morph`, ...).

    Search order: `$LATTICE_TAGGER_RESOURCES/<name>`, this package's `resources/<name>`, an
    importable `lattice_tagger` package's `resources/<name>`.  The large `base` dictionary is data
    of the reference project and is not duplicated in this repository.
    """
    candidates = []
    env = os.environ.get('LATTICE_TAGGER_RESOURCES')
    if env:
        candidates.append(os.path.join(env, name))
    here = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
    candidates.append(os.path.join(here, 'resources', name))
    try:
        import importlib.util
        spec = importlib.util.find_spec('lattice_tagger')
        if spec is not None and spec.submodule_search_locations:
            for loc in spec.submodule_search_locations:
                candidates.append(os.path.join(loc, 'resources', name))
    except (ImportError, ValueError):
        pass
    for cand in candidates:
        if os.path.isdir(cand):
            return cand
    raise FileNotFoundError(
        "dictionary resources '%s' not found; set LATTICE_TAGGER_RESOURCES to the directory that "
        "holds it (searched: %s)" % (name, ', '.join(candidates)))


class DemoMorphemeDictionary(MorphemeDictionary):
    """The 28-entry development dictionary (reference `dictionary.py:328-336`)."""

    def __init__(self):
        directory = find_resources('demo_morph')
        super().__init__(load_dictionary(directory), load_rules(os.path.join(directory, 'rules')))


class BaseMorphemeDictionary(MorphemeDictionary):
    """The Sejong-derived full dictionary (reference `dictionary.py:339-347`)."""

    def __init__(self):
        directory = find_resources('base')
        super().__init__(load_dictionary(directory), load_rules(os.path.join(directory, 'rules')))
