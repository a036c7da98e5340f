"""Drop-in alias: ``import bitblas`` -> ``bitblas_b200`` (same objects, same submodule names), so GPTQModel / vLLM /
BitNet integration code written against microsoft/BitBLAS imports unchanged."""
import importlib
import sys

import bitblas_b200 as _impl

for _name in ("cache", "quantization", "quantization.utils", "testing", "module", "ops", "ops.operator",
              "ops.general_matmul", "utils", "parallel"):
    sys.modules[f"bitblas.{_name}"] = importlib.import_module(f"bitblas_b200.{_name}")

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
__version__ = _impl.__version__
