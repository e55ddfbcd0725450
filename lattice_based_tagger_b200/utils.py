"""Small host utilities kept from the reference surface (`lattice_tagger/utils.py:6-25`)."""

import os

installpath = os.path.dirname(os.path.realpath(__file__))


def left_space_tag(sent):
    """Space-stripped characters and a 0/1 list marking syllables that start an eojeol."""
    chars = sent.replace(' ', '')
    tags = [1] + [0] * (len(chars) - 1)
    idx = 0
    for c in sent:
        if c == ' ':
            if idx < len(tags):
                tags[idx] = 1
        else:
            idx += 1
    return chars, tags
