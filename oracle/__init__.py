"""CPU oracle for the xsmm TPPs -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference``) may import this package; ``tpp_mlir_b200`` never does.
bf16 tensors are numpy ``uint16`` arrays holding the raw bits.
"""
from .pyoracle import (  # noqa: F401
    BF16,
    F32,
    TensorInit,
    bf16_to_f32,
    binary,
    brgemm,
    build,
    f32_to_bf16,
    fast_isa,
    fused_brgemm,
    fused_brgemm_amx,
    fused_brgemm_amx_grid,
    fused_brgemm_fast,
    gemm,
    has_amx,
    lib,
    num_threads,
    set_acc_mode,
    set_num_threads,
    set_vnni_factor,
    tensor_pack,
    tensor_unpack,
    unary,
    use_native,
)
