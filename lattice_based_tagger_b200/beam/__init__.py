from .beam import Sequence
from .score_funcs import BeamScoreFunction
from .score_funcs import BeamScoreFunctions
from .score_funcs import MorphemePreferenceScore
from .score_funcs import RegularizationScore
from .score_funcs import SimpleTrigramFeatureScore
from .score_funcs import WordPreferenceScore
