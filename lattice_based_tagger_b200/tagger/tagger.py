"""placeholder"""
class Tagger:
    pass
