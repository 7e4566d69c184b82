"""ctypes binding of liblnasr_b200.so (the C ABI declared in include/lnasr_b200.h).

The shared library holds every hand-written sm_100a kernel of the acoustic hot path.  There is NO CPU
fallback: if the library is missing, or no CUDA device is present, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblnasr_b200.so")

DT_I16, DT_F32, DT_F64 = 0, 1, 2

_lib = None

vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64
sz = C.c_size_t
f64 = C.c_double


class MfccTables(C.Structure):
    """lnasr_mfcc_tables (include/lnasr_b200.h)."""
    _fields_ = [
        ("frame_len", i32), ("frame_step", i32), ("fft_n", i32), ("n_mel", i32), ("n_ceps", i32),
        ("hpf_alpha", C.c_double), ("fb_stride", i32),
        ("window", vp), ("twiddle", vp), ("fb_start", vp), ("fb_len", vp), ("fb_weight", vp), ("dct_t", vp),
        ("mel_plan", vp),
    ]


class HmmBank(C.Structure):
    """lnasr_hmm_bank (include/lnasr_b200.h)."""
    _fields_ = [
        ("n_models", i32), ("N", i32), ("logA", vp), ("logpi", vp), ("max_in", i32), ("max_out", i32),
        ("pred_idx", vp), ("pred_val", vp), ("succ_idx", vp), ("succ_val", vp), ("band", i32),
    ]


class Recognizer(C.Structure):
    """lnasr_recognizer (include/lnasr_b200.h)."""
    _fields_ = [("mfcc", C.POINTER(MfccTables)), ("bank", C.POINTER(HmmBank)), ("gmm_packed", vp),
                ("gmm_kind", i32), ("n_mix", i32), ("dim", i32)]


# name -> (restype, argtypes); every symbol include/lnasr_b200.h declares
SIGNATURES = {
    "lnasr_last_error": (C.c_char_p, []),
    "lnasr_version": (C.c_int, []),
    "lnasr_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(sz)]),
    "lnasr_preemphasis": (C.c_int, [vp, C.c_int, i64, f64, vp, vp]),
    "lnasr_mfcc_num_frames": (i64, [i64, i32, i32]),
    "lnasr_mfcc_workspace_bytes": (C.c_int, [i64, i32, C.POINTER(sz)]),
    "lnasr_mfcc_mel_plan_bytes": (sz, []),
    "lnasr_mfcc_mel_plan": (C.c_int, [C.POINTER(MfccTables), vp, vp]),
    "lnasr_mfcc_batch": (C.c_int, [vp, C.c_int, vp, vp, i32, i64, C.POINTER(MfccTables), vp, i32, vp, vp, vp, sz, vp]),
    "lnasr_gmm_exact_packed_doubles": (i64, [i32, i32, i32]),
    "lnasr_gmm_pack_exact": (C.c_int, [vp, vp, vp, i32, i32, i32, vp, vp]),
    "lnasr_gmm_loglik_exact": (C.c_int, [vp, C.c_int, i32, i64, vp, i32, i32, i32, vp, C.c_int, i64, vp, vp]),
    "lnasr_gmm_loglik_full": (C.c_int, [vp, C.c_int, i32, i64, vp, vp, vp, vp, i32, i32, i32, vp, C.c_int, i64, vp, vp]),
    "lnasr_gmm_tc_packed_bytes": (C.c_int, [i32, i32, i32, C.POINTER(sz)]),
    "lnasr_gmm_pack_tc": (C.c_int, [vp, vp, vp, i32, i32, i32, vp, vp, vp]),
    "lnasr_gmm_tc_workspace_bytes": (C.c_int, [i64, C.POINTER(sz)]),
    "lnasr_gmm_loglik_tc": (C.c_int, [vp, i32, i64, vp, i32, i32, i32, vp, i64, vp, sz, vp]),
    "lnasr_gmm_tc16_packed_bytes": (C.c_int, [i32, i32, i32, C.POINTER(sz)]),
    "lnasr_gmm_pack_tc16": (C.c_int, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]),
    "lnasr_gmm_loglik_tc16": (C.c_int, [vp, i32, i64, vp, i32, i32, i32, vp, i64, vp, sz, vp]),
    "lnasr_debug_set_umma_swap": (C.c_int, [C.c_int]),
    "lnasr_debug_set_gmm_simt": (C.c_int, [C.c_int]),
    "lnasr_debug_fp64_peak_scratch_doubles": (i64, []),
    "lnasr_debug_fp64_peak": (C.c_int, [C.c_int, vp, C.POINTER(C.c_double), vp]),
    "lnasr_emit_gather": (C.c_int, [vp, vp, i64, i32, i32, vp, vp]),
    "lnasr_trans_degree": (C.c_int, [vp, i32, i32, vp, vp]),
    "lnasr_trans_pack": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "lnasr_viterbi_workspace_bytes": (sz, [i64, i32, i32, i32]),
    "lnasr_viterbi_batch": (C.c_int, [vp, C.c_int, i64, vp, i32, i64, C.POINTER(HmmBank), vp, vp, vp, vp, vp, sz, vp]),
    "lnasr_forward_batch": (C.c_int, [vp, C.c_int, i64, vp, i32, i64, C.POINTER(HmmBank), vp, vp, vp]),
    "lnasr_backward_batch": (C.c_int, [vp, C.c_int, i64, vp, i32, i64, C.POINTER(HmmBank), vp, vp]),
    "lnasr_argmax_rows": (C.c_int, [vp, i32, i32, vp, vp, vp]),
    "lnasr_recognize_workspace_bytes": (C.c_int, [i64, i32, C.POINTER(Recognizer), C.c_int, C.POINTER(sz)]),
    "lnasr_recognize_batch": (C.c_int, [vp, C.c_int, vp, vp, i32, i64, C.POINTER(Recognizer), vp, vp, vp, vp, sz, vp]),
    "lnasr_em_acc_doubles": (i64, [i32, i32, i32]),
    "lnasr_fb_workspace_bytes": (sz, [i64, i32]),
    "lnasr_fb_accumulate": (C.c_int, [vp, C.c_int, i32, vp, vp, vp, vp, i32, i64, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, sz, vp]),
    "lnasr_em_statistics": (C.c_int, [vp, C.c_int, i32, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, i32, i32, i32, vp, vp, vp]),
    "lnasr_em_statistics_workspace_bytes": (sz, [i64, i32]),
    "lnasr_em_statistics_ws": (C.c_int, [vp, C.c_int, i32, vp, vp, vp, vp, vp, i32, i64, vp, vp, vp, i32, i32, i32, vp, vp, vp, sz, vp]),
    "lnasr_hmm_em_statistics_ws": (C.c_int, [vp, vp, vp, vp, vp, i32, i64, vp, i32, i32, vp, vp, vp, sz, vp]),
    "lnasr_mstep_finalize": (C.c_int, [vp, i32, i32, i32, f64, vp, vp, vp, vp, vp, vp]),
    "lnasr_hmm_acc_doubles": (i64, [i32, i32]),
    "lnasr_hmm_fb_accumulate": (C.c_int, [vp, vp, vp, vp, i32, i64, vp, i32, i32, vp, vp, vp, vp]),
    "lnasr_hmm_mstep_finalize": (C.c_int, [vp, i32, i32, vp, vp, vp, vp]),
    "lnasr_baumwelch_xi": (C.c_int, [vp, vp, vp, vp, vp, i32, i64, i32, vp, vp, vp]),
    "lnasr_launch_count": (i64, []),
    "lnasr_vad_num_frames": (i64, [i64, i32, i32]),
    "lnasr_vad_workspace_bytes": (sz, [i64, i32, i32]),
    "lnasr_vad_ltsd": (C.c_int, [vp, C.c_int, i64, i32, i32, i32, f64, C.c_int, f64, vp, vp, vp, vp, sz, vp]),
    "lnasr_debug_set_probe_second_mma": (C.c_int, [C.c_uint32]),
    "lnasr_debug_set_band_kernel": (C.c_int, [C.c_int]),
    "lnasr_debug_set_tc16_poly": (C.c_int, [C.c_int]),
    "lnasr_debug_set_tc16_groups": (C.c_int, [C.c_int]),
    "lnasr_debug_set_em_moments_simt": (C.c_int, [C.c_int]),
    "lnasr_debug_umma_probe": (C.c_int, [vp, i32, vp, i32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_uint32, vp, vp]),
}


class LnasrError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise if it is absent -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LnasrError(
                f"{LIB_PATH} not found: build it with `python ln-asr_b200/build.py` "
                "(the acoustic path has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().lnasr_last_error().decode("utf-8", "replace")
        raise LnasrError(f"{what or 'lnasr'} failed (code {rc}): {msg}")


_torch = None


def require_cuda():
    """torch, after checking once that a CUDA device exists (the check costs microseconds per call otherwise)."""
    global _torch
    if _torch is None:
        import torch
        if not torch.cuda.is_available():
            raise LnasrError("lnasr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _torch = torch
    return _torch


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(require_cuda().cuda.current_stream().cuda_stream)
