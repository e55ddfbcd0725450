"""lattice_based_tagger_b200 — B200-native batched decode path of the lattice tagger.

Same Python surface as the reference package `lattice_tagger` for its decode path
(`Tagger` / `tag()`, the `dictionary` and `features` loaders, the score-function classes and the
trainer's weight format); dictionary lookup, lattice construction, feature scoring and beam
search run as hand-written sm_100a CUDA kernels behind the C-ABI library `liblt_b200.so`
(`include/lt_b200.h`).  There is no CPU fallback: using the tagger without the built library or
without a CUDA device raises.
"""

from .tagset import Noun, Pronoun, Number, Josa, Adjective, Verb, Eomi, Adverb, Determiner
from .tagset import Exclamation, BOS, EOS, Unk
from .utils import installpath, left_space_tag

from . import beam
from . import dictionary
from . import features
from . import tagger
from . import trainer
from .tagger import Tagger

__version__ = '0.1.0'
