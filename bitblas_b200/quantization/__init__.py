from .utils import gen_quant4, general_compress, interleave_weight  # noqa: F401
