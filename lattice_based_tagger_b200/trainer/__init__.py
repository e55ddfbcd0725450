from .params import load_params
from .params import dump_params
