from .operator import Operator, OperatorConfig, TransformKind, OptimizeStrategy  # noqa: F401
from .general_matmul import Matmul, MatmulConfig  # noqa: F401
