from .tagger import Tagger
