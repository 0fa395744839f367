"""vsiquantization_b200 -- B200-native (sm_100a) implementation of VSIQuantization's QAT fake-quantization hot path.

Host side: Python/PyTorch mirroring the reference's plugin API (registry-built UniformQuantizer / LSQQuantizer /
MinMaxObserver / LSQObserver, FuseConfig + YAML config manager, fuse_modules_unified, calibrate_qat_model /
activate_learning_qparam, reestimate_BN_stats).  Device side: hand-written CUDA kernels behind the C ABI in
include/vsiq.h (libvsiq.so).  There is no CPU fallback: importing this package without the built library fails.
"""
from . import _lib  # noqa: F401  (raises ImportError when libvsiq.so is missing)

__version__ = "0.1.0"
