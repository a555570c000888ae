"""Hot ops: native sm_100a kernels on CUDA tensors, torch reference on CPU tensors."""

from . import reference  # noqa: F401
from .functional import (  # noqa: F401
    add_rmsnorm,
    attention,
    attention_qkv,
    cross_entropy,
    embedding,
    gemm_wgather,
    gemm,
    gemm_mxfp8,
    launch_count,
    mxfp8_sf_bytes,
    quantize_mxfp8,
    linear,
    linear_mxfp8,
    linear_qkv_rope,
    linear_swiglu,
    rope_fusable,
    reset_launch_count,
    rmsnorm,
    rope_attention_qkv,
    rope_qkv,
    set_gemm_pair_mode,
    set_gemm_split_k,
    set_mxfp8_pair_mode,
    swiglu,
)
