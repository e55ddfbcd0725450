from .feature import SimpleTrigramEncoder
from .feature import WordsEncoder
from .feature import trigram_encoder
