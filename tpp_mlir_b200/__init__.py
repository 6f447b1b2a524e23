"""tpp_mlir_b200 - B200-native execution backend for tpp-mlir's xsmm dialect.

The product is ``lib/libtpp_xsmm_runner_utils.so`` (hand-written sm_100a CUDA
behind the reference's runtime/Xsmm C-ABI, see include/tpp_xsmm_abi.h). This
package holds its sources (``csrc/``), the in-tree build (``_build``), a ctypes
mirror of the dispatch/invoke surface (``xsmm``) and the replay of tpp-run's
call sequences for the benchmark workloads (``harness``).
"""
from . import _build  # noqa: F401

__all__ = ["_build", "xsmm", "harness"]


def __getattr__(name):
    # xsmm loads (and if needed builds) the CUDA library: import lazily so that
    # `import tpp_mlir_b200` itself never needs a toolchain or a GPU.
    if name in ("xsmm", "harness"):
        import importlib

        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
