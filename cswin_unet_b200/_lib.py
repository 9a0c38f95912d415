"""ctypes binding of libcswin_b200.so (C ABI declared in include/cswin_b200.h).

The product has no CPU path and no fallback: if the shared library is missing or a call fails, an
exception is raised.  `build()` compiles the library in-tree with nvcc for sm_100a (cross-compiles
without a GPU); the built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("CSWIN_LIB_PATH") or os.path.join(_HERE, "libcswin_b200.so")   # override: A/B builds of the same ABI
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "cswin_b200.h")

F32, BF16 = 0, 1
ABI_VERSION = 8

c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class LepeBranch(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("q_bs", c_int64), ("q_ts", c_int64), ("k_bs", c_int64), ("k_ts", c_int64),
        ("v_bs", c_int64), ("v_ts", c_int64),
        ("out", c_void_p), ("o_bs", c_int64), ("o_ts", c_int64),
        ("conv_w", c_void_p), ("conv_b", c_void_p), ("lse", c_void_p),
        ("C_b", c_int32), ("heads", c_int32), ("H_sp", c_int32), ("W_sp", c_int32),
    ]


class LepeBranchGrad(C.Structure):
    _fields_ = [
        ("fwd", LepeBranch),
        ("dout", c_void_p), ("do_bs", c_int64), ("do_ts", c_int64),
        ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("dq_bs", c_int64), ("dq_ts", c_int64), ("dk_bs", c_int64), ("dk_ts", c_int64),
        ("dv_bs", c_int64), ("dv_ts", c_int64),
        ("dconv_w", c_void_p), ("dconv_b", c_void_p),
    ]


class LinearArgs(C.Structure):
    _fields_ = [
        ("a", c_void_p), ("lda", c_int64), ("K1", c_int32),
        ("a2", c_void_p), ("lda2", c_int64), ("K2", c_int32),
        ("w", c_void_p), ("ldw", c_int64),
        ("bias", c_void_p),
        ("ln_gamma", c_void_p), ("ln_beta", c_void_p), ("ln_eps", c_float),
        ("residual", c_void_p), ("ldr", c_int64),
        ("sample_scale", c_void_p), ("rows_per_sample", c_int32),
        ("out", c_void_p), ("ldo", c_int64),
        ("M", c_int64), ("N", c_int32),
        ("act", c_int32), ("w_layout", c_int32),
        ("ln_stats", c_void_p), ("ln_stats_parts", c_int32), ("ln_C", c_int32),
        ("ln_colsum", c_void_p), ("bias_f32", c_void_p), ("stats_out", c_void_p),
        ("aux_out", c_void_p), ("ld_aux", c_int64),
    ]


# symbol -> (restype, argtypes); every symbol include/cswin_b200.h declares
class MlpArgs(C.Structure):
    """== cswin_mlp_args_t"""
    _fields_ = [
        ("x", c_void_p), ("ldx", c_int64), ("w1", c_void_p), ("ldw1", c_int64), ("ln_colsum", c_void_p), ("b1", c_void_p),
        ("w2", c_void_p), ("ldw2", c_int64), ("b2", c_void_p), ("ln_stats", c_void_p), ("ln_stats_parts", c_int32),
        ("ln_eps", C.c_float), ("out", c_void_p), ("ldo", c_int64), ("stats_out", c_void_p), ("M", c_int64),
        ("C", c_int32), ("hidden", c_int32),
    ]


class QkvAttnBranch(C.Structure):
    """== cswin_qkv_attn_branch_t"""
    _fields_ = [("conv_w", c_void_p), ("conv_b", c_void_p), ("heads", c_int32), ("H_sp", c_int32), ("W_sp", c_int32),
                ("reserved", c_int32)]


class QkvAttnArgs(C.Structure):
    """== cswin_qkv_attn_args_t"""
    _fields_ = [
        ("x", c_void_p), ("x_bs", c_int64), ("x_ts", c_int64), ("w", c_void_p), ("ldw", c_int64), ("bias_f32", c_void_p),
        ("ln_stats", c_void_p), ("ln_colsum", c_void_p), ("ln_stats_parts", c_int32), ("ln_eps", c_float),
        ("out", c_void_p), ("o_bs", c_int64), ("o_ts", c_int64),
        ("B", c_int32), ("reso", c_int32), ("C", c_int32), ("n_branches", c_int32),
        ("br", QkvAttnBranch * 2), ("scale", c_float), ("reserved", c_int32),
    ]


class StageBlock(C.Structure):
    """== cswin_stage_block_t"""
    _fields_ = [
        ("w_qkv", c_void_p), ("cs_qkv", c_void_p), ("b_qkv", c_void_p), ("w_proj", c_void_p), ("b_proj", c_void_p),
        ("w_fc1", c_void_p), ("cs_fc1", c_void_p), ("b_fc1", c_void_p), ("w_fc2", c_void_p), ("b_fc2", c_void_p),
        ("lepe_w", c_void_p * 2), ("lepe_b", c_void_p * 2), ("eps1", c_float), ("eps2", c_float),
    ]


class StageArgs(C.Structure):
    """== cswin_stage_args_t"""
    _fields_ = [
        ("x", c_void_p), ("stats_in", c_void_p), ("stats_in_parts", c_int32), ("n_blocks", c_int32),
        ("qkv", c_void_p), ("att", c_void_p), ("x1", c_void_p), ("hid", c_void_p),
        ("stats_x", c_void_p), ("stats_x1", c_void_p), ("ctrl", c_void_p), ("ctrl_ints", c_int64),
        ("blocks", C.POINTER(StageBlock)),
        ("B", c_int32), ("reso", c_int32), ("C", c_int32), ("hidden", c_int32), ("n_branches", c_int32), ("reserved", c_int32),
        ("heads", c_int32 * 2), ("H_sp", c_int32 * 2), ("W_sp", c_int32 * 2), ("scale", c_float), ("reserved2", c_int32),
    ]


class StagePlan(C.Structure):
    """== cswin_stage_plan_t"""
    _fields_ = [("parts_x", c_int32), ("parts_x1", c_int32), ("max_blocks", c_int32), ("reserved", c_int32), ("ctrl_ints", c_int64)]


SIGNATURES = {
    "cswin_abi_version": (c_int32, []),
    "cswin_last_error": (C.c_char_p, []),
    "cswin_launch_count": (C.c_uint64, []),
    "cswin_tc_launch_count": (C.c_uint64, []),
    "cswin_simt_fallback_count": (C.c_uint64, []),
    "cswin_debug_set_trace": (None, [c_void_p]),
    "cswin_set_option": (c_int32, [c_int32, c_int32]),
    "cswin_lepe_attention_fwd": (c_int32, [C.POINTER(LepeBranch), c_int32, c_int32, c_int32, c_float, c_int32, c_void_p]),
    "cswin_lepe_attention_bwd": (c_int32, [C.POINTER(LepeBranchGrad), c_int32, c_int32, c_int32, c_float, c_int32, c_void_p]),
    "cswin_lepe_param_grad": (c_int32, [C.POINTER(LepeBranchGrad), c_int32, c_int32, c_int32, c_int32, c_void_p, C.POINTER(c_int32)]),
    "cswin_zoom_cubic_fwd": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32,
                                       c_int32, c_void_p]),
    "cswin_zoom_nearest_u8": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p]),
    "cswin_layernorm_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                      c_float, c_void_p, c_void_p, c_int32, c_void_p]),
    "cswin_linear_fwd": (c_int32, [C.POINTER(LinearArgs), c_int32, c_void_p]),
    "cswin_carafe_head_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                        c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "cswin_seg_loss_fwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_int32, c_int64, c_void_p]),
    "cswin_seg_loss_bwd": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int64, c_int32,
                                     c_int64, c_void_p]),
    "cswin_sgd_momentum_step": (c_int32, [c_void_p, c_int32, c_void_p, c_float, c_float, c_void_p]),
    "cswin_mlp_fwd": (c_int32, [C.POINTER(MlpArgs), c_int32, c_void_p]),
    "cswin_mlp_stats_parts": (c_int32, [c_int32, c_int32]),
    "cswin_qkv_lepe_attention_fwd": (c_int32, [C.POINTER(QkvAttnArgs), c_int32, c_void_p]),
    "cswin_qkv_lepe_attention_supported": (c_int32, [c_int32, c_int32, c_int32, C.POINTER(c_int32), C.POINTER(c_int32),
                                                     C.POINTER(c_int32)]),
    "cswin_linear_stats_parts": (c_int32, [c_int64, c_int32, c_int32, c_int32]),
    "cswin_stage_plan": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_int32, C.POINTER(c_int32), C.POINTER(c_int32),
                                   C.POINTER(c_int32), C.POINTER(StagePlan)]),
    "cswin_stage_fwd": (c_int32, [C.POINTER(StageArgs), c_int32, c_void_p]),
    "cswin_stem_fwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int32, c_int32,
                                 c_int32, c_int32, c_void_p, C.POINTER(c_int32)]),
    "cswin_layernorm_stats_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32, c_float,
                                            c_void_p, c_int32, c_void_p]),
    "cswin_row_stats": (c_int32, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int32, c_void_p]),
    "cswin_im2col_tokens": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64] + [c_int32] * 9 + [c_void_p]),
    "cswin_conv_tokens_fwd": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64] + [c_int32] * 10 +
                              [c_void_p, C.POINTER(c_int32)]),
    "cswin_im2col_nchw": (c_int32, [c_void_p, c_int32, c_void_p, c_int64] + [c_int32] * 9 + [c_void_p]),
    "cswin_act_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "cswin_act_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_int64, c_int32,
                                c_int32, c_int32, c_void_p]),
    "cswin_linear_wgrad": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32,
                                     c_int32, c_int32, c_void_p]),
    "cswin_layernorm_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "cswin_col2im_tokens": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int64] + [c_int32] * 9 + [c_void_p]),
    "cswin_carafe_reassemble_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int64, c_int64,
                                              c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]
                                    + [c_int32] * 6 + [c_void_p]),
    "cswin_carafe_head_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p]
                              + [c_int32] * 6 + [c_void_p]),
    "cswin_carafe_reassemble_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64]
                                    + [c_int32] * 8 + [c_void_p]),
}

_lib = None
_lock = threading.Lock()


class CswinError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libcswin_b200.so in-tree for sm_100a (make -C csrc). Returns the library path."""
    proc = subprocess.run(["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise CswinError("building libcswin_b200.so failed (see output above)")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises CswinError if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise CswinError(
                    f"{LIB_PATH} is missing: the CUDA extension is the only execution path of cswin_unet_b200. "
                    "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C cswin_unet_b200/csrc`.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)     # AttributeError if the .so does not export a declared symbol
                fn.restype, fn.argtypes = res, args
            got = handle.cswin_abi_version()
            if got != ABI_VERSION:
                raise CswinError(f"libcswin_b200.so ABI version {got} != binding version {ABI_VERSION}; rebuild")
            _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().cswin_last_error()
        raise CswinError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


OPT_GEMM_SMEM_CAP_KB = 1


def set_option(option: int, value: int) -> None:
    check(lib().cswin_set_option(option, value), "cswin_set_option")


def launch_count() -> int:
    return int(lib().cswin_launch_count())


def simt_fallback_count() -> int:
    """bf16 calls that fell outside a tcgen05 kernel's envelope and ran on the general SIMT kernels (loud: warned once per op)."""
    return int(lib().cswin_simt_fallback_count())


def tc_launch_count() -> int:
    """Launches of tcgen05 / TMEM / TMA kernels so far (subset of launch_count())."""
    return int(lib().cswin_tc_launch_count())
