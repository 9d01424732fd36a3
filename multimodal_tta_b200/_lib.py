"""ctypes binding of the C-ABI in include/tta_b200.h (libtta_b200.so, built in-tree by nvcc).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised -- the product path never routes through PyTorch ops or the oracle.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# TTA_LIB selects a differently-built variant of the library (kernel A/B experiments inside one process launch)
LIB_PATH = os.environ.get("TTA_LIB") or os.path.join(_HERE, "libtta_b200.so")
CSRC = os.path.join(_HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--threads", "0", "-split-compile", "0"]

TTA_F16, TTA_BF16, TTA_F16_HI = 0, 1, 2


def build_library(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """Compile every csrc/*.cu into libtta_b200.so for sm_100a (cross-compiles without a GPU).
    ``out`` / ``defines``: build a variant (e.g. ``-DTTA_PDL_LATE``) next to the default library."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + sorted(glob.glob(os.path.join(CSRC, "*.cuh")))
    target = out or os.path.join(_HERE, "libtta_b200.so")
    if not force and os.path.exists(target):
        newest = max(os.path.getmtime(p) for p in deps)
        if os.path.getmtime(target) >= newest:
            return target
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", target, *srcs]
    if verbose:
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    return target


def _has_libcuda() -> bool:
    # the driver API entry points (cuTensorMapEncodeTiled) are resolved at run time through
    # cudaGetDriverEntryPoint, so libcuda is never needed at link time.
    return False


_lib = None

P, I, L, F = c_void_p, c_int, c_longlong, c_float

_SIGNATURES = {
    "tta_last_error": (c_char_p, []),
    "tta_abi_version": (I, []),
    "tta_device_sm": (I, []),
    "tta_norm_workspace_floats": (L, [I, I, L]),
    "tta_norm_stats": (I, [P, L, I, I, L, I, F, P, P, P, I, P]),
    "tta_norm_apply": (I, [P, L, I, I, L, P, P, P, P, I, I, P, P, L, P, P, L, I, P, I, I, F, P, P, L, I, P]),
    "tta_norm_stats_finalize": (I, [P, I, I, I, L, I, F, P, P, P]),
    "tta_norm_bwd_reduce": (I, [P, L, P, L, P, L, I, I, I, L, P, P, P, P, I, I, P, P, P, P, I, I, P]),
    "tta_norm_bwd_apply": (I, [P, L, P, L, P, L, I, I, L, P, P, P, P, I, I, P, P, P, L, P, P, L, I, P, I, P, P, I, P]),
    "tta_norm_small_supported": (I, [I, L, I]),
    "tta_norm_fwd_small": (I, [P, L, I, I, L, F, P, P, P, P, I, I, P, P, L, P, P, L, I, P, P, L, I, P]),
    "tta_norm_bwd_small": (I, [P, L, P, L, P, L, I, I, I, L, P, P, P, P, I, P, P, P, P, P, L, P, P, L, I, I, P, P]),
    "tta_split_f32": (I, [P, L, P, L, I, I, L, P, P, L, I, P]),
    "tta_gather_pack": (I, [P, I, I, I, I, I, P, P, I, I, I, I, P, P, L, I, I, P]),
    "tta_gather_pack_norm": (I, [P, I, I, I, I, I, P, P, P, I, I, I, I, P, P, L, I, I, P]),
    "tta_gather_pack_norm_f16": (I, [P, I, I, I, I, I, P, P, P, I, I, I, I, P, P, L, I, I, P]),
    "tta_head_entropy_blocks": (I, [I, L]),
    "tta_head_entropy": (I, [P, L, I, I, L, I, F, F, I, P, P, P, P, L, P, P, P]),
    "tta_head_fused_supported": (I, [I, I, I, I]),
    "tta_head_fused_tiles": (I, [I, I, I]),
    "tta_head_fused_workspace_floats": (L, [I, I, I, I]),
    "tta_head_fused_fwd": (I, [P, L, I, I, I, I, I, I, P, P, P, P, I, P, P, I, F, F, P, P, P, P, P, P]),
    "tta_head_fused_bwd": (I, [P, I, I, I, I, I, P, P, L, I, P, P, P, P, I, I, P, L, P, P, P, P, P]),
    "tta_norm_bwd_apply_c4": (I, [P, L, P, L, I, I, L, P, P, P, P, I, P, P, P, L, I, I, I, P]),
    "tta_adam_step": (I, [P, P, P, P, I, F, F, F, F, F, P, P]),
    "tta_sw_blend": (I, [P, I, I, I, I, I, P, P, P, P, P, F, P, P, I, I, I, I, P]),
    "tta_sw_normalise": (I, [P, P, I, I, L, P, P]),
    "tta_dice_counts": (I, [P, P, I, L, F, P, P]),
    "tta_intensity_workspace_bytes": (L, [I, I, L]),
    "tta_intensity_stats": (I, [P, I, I, L, P, I, P, P, P]),
    "tta_intensity_apply": (I, [P, P, I, I, L, P, P]),
    "tta_conv_simt": (I, [P, P, L, I, I, I, I, I, I, P, P, P, L, I, I, I, I, I, I, I, I, P]),
    "tta_conv_small_supported": (I, [I, I, I, I]),
    "tta_conv_small": (I, [P, P, L, I, I, I, I, I, I, P, P, P, L, I, I, P]),
    "tta_conv_tc_supported": (I, [I, I, I, I, I]),
    "tta_conv_tc_ntile": (I, [I, I, I, I, I]),
    "tta_conv_tc_stacked": (I, [I, I, I, I, I, I]),
    "tta_conv_tc_t2s": (I, [I, I, I, I, I, I]),
    "tta_conv_tc_s2pair": (I, [I, I, I, I]),
    "tta_conv_tc_s2c4": (I, [I, I, I, I]),
    "tta_conv_tc_gmax": (I, [I, I, I]),
    "tta_conv_tc_ngroups": (I, [I, I, I]),
    "tta_conv_tc_packed_bytes": (L, [I, I, I, I, I]),
    "tta_conv_tc": (I, [P, P, L, I, I, I, I, I, I, P, P, P, L, I, I, I, I, I, I, I, I, I, P, I, P]),
    "tta_conv_tc_bwd_norm": (I, [P, P, L, I, I, I, I, I, I, P, P, L, I, I, I, I, I, I, I, I, I, P, I, P]),
    "tta_norm_bwd_finalize": (I, [P, I, I, I, I, I, P, P, P, P]),
    "tta_conv_wgrad": (I, [P, P, L, I, I, I, I, P, P, L, I, I, I, I, I, I, I, I, I, I, I, F, P, I, I, P, P]),
    "tta_conv_wgrad_tc_supported": (I, [I, I, I, I, I, I]),
    "tta_conv_wgrad_tc": (I, [P, P, L, I, I, I, I, P, L, I, I, I, I, I, I, I, I, I, F, P, I, I, P, I, P]),
    "tta_pack_grad": (I, [P, I, I, L, F, P, P, L, I, P]),
    "tta_repack_weights": (I, [P, P, P, L, P, I, P]),
    "tta_bias_grad": (I, [P, P, L, I, I, I, L, F, P, I, P, P]),
    "tta_plan_create": (I, [P]),
    "tta_plan_destroy": (I, [P]),
    "tta_plan_begin": (I, [P, I]),
    "tta_plan_end": (I, []),
    "tta_plan_num_launches": (I, [P, I]),
    "tta_plan_run": (I, [P, I, P]),
    "tta_step": (I, [P, P]),
    "tta_workspace_bytes": (L, [I, I, L]),
    "tta_mean_planes": (I, [P, P, P, I, I, I, L, F, P, P, L, I, P]),
    "tta_sum_f32": (I, [P, P, I, I, I, I, L, F, P, L, I, P]),
    "tta_upsample_fwd": (I, [P, L, I, I, I, I, I, I, I, I, P, P, L, I, P]),
    "tta_upsample_bwd": (I, [P, L, I, I, I, I, I, I, I, I, P, P, L, I, P]),
    "tta_conv_tc_query": (I, [I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, P, P, P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "multimodal_tta_b200 has no CPU or PyTorch fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().tta_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libtta_b200 {what} failed (status {rc}): {msg}")
