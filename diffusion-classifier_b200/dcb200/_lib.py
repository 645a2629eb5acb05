"""ctypes binding of libdcb200.so -- the C ABI declared in include/dcb200.h.

There is no fallback: if the shared library is missing or a call fails, this raises.  PyTorch is used by the
callers only for device memory (``tensor.data_ptr()``), streams and torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCB_LIB") or os.path.join(HERE, "libdcb200.so")    # DCB_LIB: experiment builds (tools/ only)

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_GELU_TANH, ACT_GEGLU = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
MAX_SEGS = 12

c_void_p, c_int, c_i32, c_i64, c_u64, c_float = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_uint64, C.c_float


class Seg(C.Structure):
    _fields_ = [("src", c_void_p), ("C", c_i32), ("H", c_i32), ("W", c_i32), ("c_off", c_i32), ("kc", c_i32),
                ("dy", c_i32), ("dx", c_i32), ("stride", c_i32), ("nb_div", c_i32), ("_r1", c_i32)]


class GemmDesc(C.Structure):
    _fields_ = [("dtype", c_i32), ("engine", c_i32), ("NB", c_i32), ("OH", c_i32), ("OW", c_i32), ("N", c_i32),
                ("nseg", c_i32), ("_r0", c_i32), ("seg", Seg * MAX_SEGS),
                ("W", c_void_p), ("bias", c_void_p), ("rowvec", c_void_p), ("rowvec_idx", c_void_p),
                ("gate", c_void_p), ("residual", c_void_p), ("res_idx", c_void_p), ("out", c_void_p),
                ("mse_target", c_void_p), ("mse_scale", c_void_p), ("mse_part", c_void_p), ("gn_part", c_void_p),
                ("rowvec_ld", c_i32), ("gate_ld", c_i32), ("rows_per_group", c_i32), ("act", c_i32),
                ("act_post", c_i32), ("res_ld", c_i32), ("res_mod", c_i32), ("res_dtype", c_i32),
                ("out_ld", c_i32), ("out_dtype", c_i32), ("mse_div", c_i32), ("mse_ld", c_i32), ("up_phase", c_i32),
                ("xf_silu", c_i32), ("xf_a", c_void_p), ("xf_b", c_void_p), ("xf_src1", c_void_p), ("xf_c1", c_i32),
                ("xf_div1", c_i32), ("attn_norms", c_void_p), ("attn_heads", c_i32), ("attn_tok", c_i32)]


_PROTOS = {
    "dcb_version": (c_int, []),
    "dcb_last_error": (C.c_char_p, []),
    "dcb_launch_count": (c_i64, []),
    "dcb_note_graph_replay": (None, [c_i64]),
    "dcb_set_knobs": (None, [C.c_uint32]),
    "dcb_get_knobs": (C.c_uint32, []),
    "dcb_prologue": (c_int, [c_int, c_int, c_void_p, c_void_p, c_u64, c_i64, c_void_p, c_void_p, c_void_p, c_int, c_int,
                             c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dcb_ddpm_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_u64, c_i64, c_int,
                              c_int, c_int, c_int, c_void_p, c_void_p]),
    "dcb_timestep_embed": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "dcb_gemm": (c_int, [C.POINTER(GemmDesc), c_void_p]),
    "dcb_struct_size": (c_int, [c_int]),
    "dcb_gemm_mse_layout": (c_int, [C.POINTER(GemmDesc), C.POINTER(c_i32), C.POINTER(c_i32)]),
    "dcb_gemm_gn_layout": (c_int, [C.POINTER(GemmDesc), C.POINTER(c_i32)]),
    "dcb_gemm_xf_layout": (c_int, [C.POINTER(GemmDesc), C.POINTER(c_i32)]),
    "dcb_groupnorm_coef_from_tiles": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                              c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "dcb_groupnorm_stats_from_tiles": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                               c_void_p, c_void_p]),
    "dcb_mse_finalize": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dcb_eps_mse": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_i64, c_void_p, c_int, c_void_p]),
    "dcb_groupnorm_stats": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p]),
    "dcb_groupnorm_apply": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p]),
    "dcb_groupnorm_stats_div": (c_int, [c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p]),
    "dcb_groupnorm_apply_div": (c_int, [c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p]),
    "dcb_groupnorm_fused": (c_int, [c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_float, c_int, c_void_p, c_void_p]),
    "dcb_expand_samples": (c_int, [c_int, c_void_p, c_int, c_int, c_i64, c_void_p, c_void_p]),
    "dcb_layernorm": (c_int, [c_int, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                              c_int, c_void_p, c_void_p]),
    "dcb_attention": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                              c_void_p, c_int, c_void_p]),
    "dcb_attention_ws": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                 c_void_p, c_int, c_void_p, c_void_p]),
    "dcb_haar_dwt": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dcb_haar_idwt": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dcb_upsample2x": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dcb_nhwc_to_nchw": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dcb_unpatchify": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dcb_cast_f32": (c_int, [c_int, c_void_p, c_i64, c_void_p, c_void_p]),
    "dcb_pack_conv": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dcb_pack_geglu": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dcb_pack_upsample": (c_int, [c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "dcb_pack_rows": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
}

EXPORTS = tuple(_PROTOS)
_lib = None

# DCB_KNOB_* of include/dcb200.h.  The environment variables of the same name (DCB_NO_TC2=1 ...) are read ONCE, when the
# library is loaded; tests switch at run time with ``knob()``.
ATTN_NORMS_READY = 0x200   # dcb_attention_ws dtype flag (include/dcb200.h)
KNOBS = {"NO_TC2": 1, "TC2_NO_HALO": 2, "TC2_NO_YHALO": 4, "NO_TC2_MSE": 8, "TC2_WIDE": 16, "TC_DIRECT_EPILOGUE": 32,
         "ATTN_NO_TC": 64, "ATTN_NO_FAST": 128, "NO_TC3": 256, "TC2X_NO_PAIR": 512}


class DcbError(RuntimeError):
    pass


def lib():
    """Load libdcb200.so (once).  Fails loudly -- there is no CPU / torch fallback for the hot path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -m dcb200.build, or "
                f"__graft_entry__.build()).  dcb200 has no fallback path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)  # AttributeError here == ABI drift
            fn.restype, fn.argtypes = res, args
        if L.dcb_struct_size(0) != C.sizeof(Seg) or L.dcb_struct_size(1) != C.sizeof(GemmDesc):
            raise ImportError("dcb200: ctypes struct layout does not match libdcb200.so (rebuild the extension)")
        mask = 0
        for name, bit in KNOBS.items():
            if os.environ.get("DCB_" + name, "0") not in ("", "0"):
                mask |= bit
        L.dcb_set_knobs(mask)
        _lib = L
    return _lib


class knob:
    """``with knob("NO_TC2", "TC2_NO_YHALO"): ...`` -- run the enclosed launches on an alternative (bit-identical) kernel."""

    def __init__(self, *names):
        self.mask = 0
        for n in names:
            self.mask |= KNOBS[n]

    def __enter__(self):
        self.old = lib().dcb_get_knobs()
        lib().dcb_set_knobs(self.old | self.mask)
        return self

    def __exit__(self, *exc):
        lib().dcb_set_knobs(self.old)
        return False


def check(rc, what=""):
    if rc != 0:
        msg = lib().dcb_last_error().decode("utf-8", "replace")
        raise DcbError(f"dcb200 {what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().dcb_launch_count())
