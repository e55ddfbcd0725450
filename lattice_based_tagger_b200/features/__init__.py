from .feature import SimpleTrigramEncoder
from .feature import WordsEncoder
from .feature import trigram_encoder
from .utils import scan_dictionary
from .utils import scan_features
