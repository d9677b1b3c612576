"""rabbitsalign_b200 -- B200-native engine for RabbitSAlign's extension hot path.

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of include/rsa_ext.h, built into
librsa_ext.so), `ext.py` (ctypes binding + the host-side mirror of the reference's `solve_ssw_on_gpu`
batch interface) and `workload.py` (synthetic pair generators for tests and bench.py).
"""
from .ext import ExtensionEngine, ExtensionError, GasalTmpRes, load_library  # noqa: F401
