"""ctypes binding of include/bcnf_b200.h (the C-ABI shared library ``libbcnf_b200.so``).

There is no fallback: if the library is missing or does not load, importing this module's
``lib()`` raises, and every compute entry point of the package raises with it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
# BCNF_B200_LIB: load another build of the library (A/B timing of two kernel versions on the same box)
LIB_PATH = os.environ.get("BCNF_B200_LIB") or os.path.join(HERE, "libbcnf_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "bcnf_abi.cu"), os.path.join(HERE, "csrc", "trf_abi.cu")]
HEADERS = sorted(os.path.join(HERE, "csrc", n) for n in os.listdir(os.path.join(HERE, "csrc"))
                 if n.endswith((".cuh", ".h"))) + [os.path.join(REPO, "include", "bcnf_b200.h")]

MAX_HIDDEN_LAYERS = 8
OP_ACTNORM, OP_COUPLING, OP_ORTHO = 0, 1, 2
PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2
KERNEL_NAMES = {0: "rowthread", 1: "tiled", 2: "tcgen05"}

EXPORTS = ["bcnf_abi_version", "bcnf_last_error", "bcnf_flow_create", "bcnf_flow_destroy",
           "bcnf_flow_info", "bcnf_flow_debug_words", "bcnf_flow_set_params", "bcnf_cond_project", "bcnf_flow_forward",
           "bcnf_flow_inverse", "bcnf_flow_sample", "bcnf_flow_sample_ranks", "bcnf_resimulate", "bcnf_train_gemm", "bcnf_train_colsum", "bcnf_train_dropout_mask",
           "bcnf_train_set_gemm_mode", "bcnf_train_gemm_trace", "bcnf_train_pre", "bcnf_train_post",
           "bcnf_train_post_bwd", "bcnf_train_pre_bwd", "bcnf_img_pack", "bcnf_gemm_img", "bcnf_gemm_img_gelu", "bcnf_gemm_img_set_trace", "bcnf_lstm_step",
           "bcnf_adam_flat", "bcnf_train_nll", "bcnf_trf_embed", "bcnf_trf_attention", "bcnf_trf_add_layernorm",
           "bcnf_trf_attention_bwd", "bcnf_trf_gelu", "bcnf_trf_ln_param_grad"]
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU_DROP, EPI_DGELU_DROP = 0, 1, 2, 3


class FlowDesc(C.Structure):
    _fields_ = [("size", C.c_int32), ("n_conditions", C.c_int32), ("n_hidden", C.c_int32),
                ("hidden", C.c_int32 * MAX_HIDDEN_LAYERS), ("two_way", C.c_int32), ("n_ops", C.c_int32),
                ("precision", C.c_int32), ("device", C.c_int32)]


FloatPP = C.POINTER(C.c_void_p)


class OpParams(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("scale", C.c_void_p), ("bias", C.c_void_p),
                ("q", C.c_void_p), ("w_a", FloatPP), ("b_a", FloatPP), ("w_b", FloatPP), ("b_b", FloatPP)]


class FlowInfo(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("proj_width", C.c_int32), ("n_half_couplings", C.c_int32),
                ("rows_per_cta", C.c_int32), ("packed_bytes", C.c_int64), ("macs_per_row", C.c_int64),
                ("macs_per_instance", C.c_int64)]


class GemmArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("M", C.c_int32), ("N", C.c_int32),
                ("K", C.c_int32), ("as0", C.c_int64), ("as1", C.c_int64), ("bs0", C.c_int64), ("bs1", C.c_int64),
                ("cs0", C.c_int64), ("beta", C.c_float), ("epilogue", C.c_int32), ("bias", C.c_void_p),
                ("save", C.c_void_p), ("saved", C.c_void_p), ("seed", C.c_uint64), ("layer_uid", C.c_uint32),
                ("p_drop", C.c_float), ("seed_ptr", C.c_void_p), ("colsum", C.c_void_p), ("ws", C.c_void_p),
                ("counters", C.c_void_p), ("ws_floats", C.c_int64), ("n_counters", C.c_int32), ("split_k", C.c_int32),
                ("a_img", C.c_void_p), ("a_plane", C.c_int64), ("a_rpad", C.c_int32), ("img_mn", C.c_int32),
                ("b_img", C.c_void_p), ("b_plane", C.c_int64), ("b_rpad", C.c_int32), ("pad1", C.c_int32),
                ("c_img", C.c_void_p), ("c_plane", C.c_int64), ("c_rpad", C.c_int32), ("pad2", C.c_int32)]


class ImgPackDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("s_row", C.c_int64), ("s_k", C.c_int64), ("rows", C.c_int32), ("k", C.c_int32),
                ("dst", C.c_void_p), ("plane", C.c_int64), ("rpad", C.c_int32), ("chunks", C.c_int32)]


GLUE_ORTHO, GLUE_ACTNORM = 0, 1


class GlueOp(C.Structure):
    _fields_ = [("type", C.c_int32), ("p0", C.c_void_p), ("p1", C.c_void_p), ("save", C.c_void_p), ("g0", C.c_void_p),
                ("g1", C.c_void_p)]


class TrainPreArgs(C.Structure):
    _fields_ = [("y", C.c_void_p), ("y_pitch", C.c_int64), ("B", C.c_int32), ("D", C.c_int32), ("src0", C.c_int32),
                ("din", C.c_int32), ("W1", C.c_void_p), ("w1_pitch", C.c_int64), ("P", C.c_void_p), ("p_pitch", C.c_int64),
                ("H", C.c_int32), ("pre", C.c_void_p), ("act", C.c_void_p), ("pitch", C.c_int64), ("seed", C.c_uint64),
                ("layer_uid", C.c_uint32), ("p_drop", C.c_float), ("seed_ptr", C.c_void_p), ("act_img", C.c_void_p),
                ("img_plane", C.c_int64), ("img_rpad", C.c_int32)]


class TrainPostArgs(C.Structure):
    _fields_ = [("a", C.c_void_p), ("a_pitch", C.c_int64), ("Wout", C.c_void_p), ("bout", C.c_void_p), ("B", C.c_int32),
                ("D", C.c_int32), ("H", C.c_int32), ("dst0", C.c_int32), ("dout", C.c_int32), ("y_in", C.c_void_p),
                ("y_out", C.c_void_p), ("ld", C.c_void_p), ("ls_save", C.c_void_p), ("ydst_save", C.c_void_p),
                ("n_ops", C.c_int32), ("ops", GlueOp * 4)]


class TrainPostBwdArgs(C.Structure):
    _fields_ = [("dz_in", C.c_void_p), ("dz_out", C.c_void_p), ("dld", C.c_void_p), ("B", C.c_int32), ("D", C.c_int32),
                ("H", C.c_int32), ("dst0", C.c_int32), ("dout", C.c_int32), ("ls_save", C.c_void_p),
                ("ydst_save", C.c_void_p), ("Wout", C.c_void_p), ("pre", C.c_void_p), ("pitch", C.c_int64),
                ("d_o", C.c_void_p), ("d_pre", C.c_void_p), ("seed", C.c_uint64), ("layer_uid", C.c_uint32),
                ("p_drop", C.c_float), ("seed_ptr", C.c_void_p), ("n_ops", C.c_int32), ("ops", GlueOp * 4),
                ("dpre_img", C.c_void_p), ("img_plane", C.c_int64), ("img_rpad", C.c_int32)]


class TrainPreBwdArgs(C.Structure):
    _fields_ = [("d_pre", C.c_void_p), ("pitch", C.c_int64), ("W1", C.c_void_p), ("w1_pitch", C.c_int64), ("B", C.c_int32),
                ("D", C.c_int32), ("H", C.c_int32), ("src0", C.c_int32), ("din", C.c_int32), ("dz", C.c_void_p)]


class LstmStep(C.Structure):
    _fields_ = [("a_hi", C.c_void_p * 16), ("a_lo", C.c_void_p * 16), ("b_img", C.c_void_p), ("b_plane", C.c_int64),
                ("bias", C.c_void_p), ("cell", C.c_void_p), ("hsum", C.c_void_p), ("state_rows", C.c_int64),
                ("n_chunks", C.c_int32), ("a_rpad", C.c_int32), ("b_rpad", C.c_int32), ("M", C.c_int32),
                ("h_hi", C.c_void_p * 4), ("h_lo", C.c_void_p * 4), ("N", C.c_int32), ("passes", C.c_int32),
                ("pad0", C.c_int32), ("pad1", C.c_int32)]


def nvcc_command(out_path: str = LIB_PATH) -> list[str]:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # (--threads: the two translation units compile side by side)
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "2",
            "-shared", "-Xcompiler", "-fPIC", "-o", out_path] + SOURCES + ["-ldl"]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``libbcnf_b200.so`` (in-tree)."""
    if force or needs_build():
        cmd = nvcc_command()
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it is absent -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"bcnf_b200: {LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the coupling stack.")
    L = C.CDLL(LIB_PATH)
    L.bcnf_abi_version.restype = C.c_int
    L.bcnf_last_error.restype = C.c_char_p
    L.bcnf_flow_create.argtypes = [C.POINTER(FlowDesc), C.POINTER(C.c_int32), C.POINTER(C.c_void_p)]
    L.bcnf_flow_destroy.argtypes = [C.c_void_p]
    L.bcnf_flow_info.argtypes = [C.c_void_p, C.POINTER(FlowInfo)]
    L.bcnf_flow_debug_words.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.bcnf_flow_set_params.argtypes = [C.c_void_p, C.POINTER(OpParams), C.c_void_p]
    L.bcnf_cond_project.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    flow_args = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                 C.c_void_p, C.c_void_p]
    L.bcnf_flow_forward.argtypes = flow_args
    L.bcnf_flow_inverse.argtypes = flow_args
    L.bcnf_resimulate.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_int32, C.c_void_p]
    L.bcnf_flow_sample.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    L.bcnf_flow_sample_ranks.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_float, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bcnf_train_gemm.argtypes = [C.POINTER(GemmArgs), C.c_int32, C.c_void_p]
    L.bcnf_adam_flat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_void_p]
    L.bcnf_train_nll.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int32, C.c_void_p]
    L.bcnf_train_set_gemm_mode.argtypes = [C.c_int32]
    L.bcnf_gemm_img.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.bcnf_gemm_img_set_trace.argtypes = [C.c_void_p]
    L.bcnf_lstm_step.argtypes = [C.POINTER(LstmStep), C.c_int32, C.c_void_p]
    L.bcnf_gemm_img_gelu.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.bcnf_trf_embed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    L.bcnf_trf_attention.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_int32, C.c_int32, C.c_void_p]
    L.bcnf_trf_attention_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_int32, C.c_void_p]
    L.bcnf_trf_add_layernorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64,
                                         C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_int32, C.c_int32, C.c_void_p]
    L.bcnf_trf_ln_param_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_void_p]
    L.bcnf_trf_gelu.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                C.c_void_p]
    L.bcnf_img_pack.argtypes = [C.POINTER(ImgPackDesc), C.c_int32, C.c_int32, C.c_void_p]
    L.bcnf_train_pre.argtypes = [C.POINTER(TrainPreArgs), C.c_int32, C.c_void_p]
    L.bcnf_train_post.argtypes = [C.POINTER(TrainPostArgs), C.c_int32, C.c_void_p]
    L.bcnf_train_post_bwd.argtypes = [C.POINTER(TrainPostBwdArgs), C.c_int32, C.c_void_p]
    L.bcnf_train_pre_bwd.argtypes = [C.POINTER(TrainPreBwdArgs), C.c_int32, C.c_void_p]
    L.bcnf_train_gemm_trace.argtypes = [C.POINTER(GemmArgs), C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]
    L.bcnf_train_colsum.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_float, C.c_int32, C.c_void_p]
    L.bcnf_train_dropout_mask.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_float,
                                          C.c_void_p, C.c_int32, C.c_void_p]
    for name in EXPORTS:
        getattr(L, name).restype = C.c_char_p if name == "bcnf_last_error" else C.c_int
    _lib = L
    return L


class BcnfError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().bcnf_last_error()
        msg = msg.decode() if msg else ""
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise BcnfError(f"{what} failed with code {rc}: {msg}")
