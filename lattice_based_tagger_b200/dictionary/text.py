"""Tagged-text helpers (host side): gold annotations -> `Word` lists.

Counterparts of the reference's `str_to_morphtag`, `text_to_words` and `flatten_words`
(`dictionary/dictionary.py:11-167`).  They are not on the decode path; the feature scanner
(`features/utils.py`) uses them to turn a tagged corpus into the `feature_dic` whose weights the
GPU feature table holds.
"""

from ..tagset import BOS, EOS
from ..utils import left_space_tag
from .dictionary import Word


def str_to_morphtag(word):
    """'이/Adjective+ㅂ니다/Eomi' -> [['이', 'Adjective'], ['ㅂ니다', 'Eomi']]"""
    return [piece.split('/', 1) for piece in word.split('+')]


def text_to_words(word_text, morph_text, sent=None):
    """Gold annotation -> [BOS, Word..., EOS].

    Eojeols are separated by two spaces, words inside an eojeol by one, the two morphemes of a
    conjugated word by '+'.  `is_l` marks the first word of every eojeol.
    """
    eojeol_words = word_text.split('  ')
    eojeol_morphs = morph_text.split('  ')
    if len(eojeol_words) != len(eojeol_morphs):
        raise ValueError('Different length of eojeols in (word_text=%d, morph_text=%d)'
                         % (len(eojeol_words), len(eojeol_morphs)))
    if sent is None:
        sent = ' '.join(eojeol.replace(' ', '') for eojeol in eojeol_words)
    _, left = left_space_tag(sent)

    words = [Word(BOS, BOS, None, BOS, None, 0, 0, 0, False)]
    begin = 0
    for surface_part, morph_part in zip(eojeol_words, eojeol_morphs):
        for surface, analysis in zip(surface_part.split(), morph_part.split()):
            morphtags = str_to_morphtag(analysis)
            if len(morphtags) > 2:
                raise ValueError('Word (%s) consists of three or more morphemes' % surface)
            n = len(surface)
            morph0, tag0 = morphtags[0]
            morph1, tag1 = morphtags[1] if len(morphtags) == 2 else (None, None)
            words.append(Word(surface, morph0, morph1, tag0, tag1, n, begin, begin + n, left[begin] == 1))
            begin += n
    words.append(Word(EOS, EOS, None, EOS, None, 0, begin, begin, False))
    return words


def flatten_words(words):
    """Split every two-morpheme word into two single-morpheme words.

    The first part spans at most its morpheme's length; an eomi that starts with a compatibility
    jamo (ㄱ..ㅎ) is one syllable shorter on the surface than its string.
    """
    flat = []
    for word in words:
        if word.tag1 is None:
            flat.append(word)
            continue
        len0, len1 = len(word.morph0), len(word.morph1)
        middle = min(word.e, word.b + len0)
        if 'ㄱ' <= word.morph1[0] <= 'ㅎ':
            len1 -= 1
        flat.append(Word(word.morph0, word.morph0, None, word.tag0, None, len0, word.b, middle, word.is_l))
        flat.append(Word(word.morph1, word.morph1, None, word.tag1, None, len1, middle, word.e, False))
    return flat
