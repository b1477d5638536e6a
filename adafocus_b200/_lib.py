"""ctypes binding of include/adafocus_b200.h.

The product path has no fallback: if the shared library is missing or a CUDA device is absent, loading / context
creation raises.  Nothing here imports oracle/.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# AF_LIB_PATH: load another build of the same ABI (A/B timing of kernel changes); default is the in-tree library
LIB_PATH = os.environ.get("AF_LIB_PATH") or os.path.join(HERE, "lib", "libadafocus_b200.so")

AF_ACT_NONE, AF_ACT_RELU, AF_ACT_RELU6 = 0, 1, 2


class AfError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """struct af_conv_desc"""
    _fields_ = [
        ("in_", c_void_p), ("w", c_void_p), ("scale", c_void_p), ("bias", c_void_p), ("residual", c_void_p),
        ("out", c_void_p),
        ("n", c_int32), ("h", c_int32), ("w_", c_int32), ("cin", c_int32), ("cout", c_int32),
        ("kh", c_int32), ("kw", c_int32), ("stride", c_int32), ("pad", c_int32),
        ("block_n", c_int32), ("act", c_int32), ("out_f32", c_int32),
        ("in_stride", c_int64), ("out_stride", c_int64), ("res_stride", c_int64),
        ("in_row_stride", c_int64), ("in_img_stride", c_int64),
        ("tsm_t", c_int32), ("tsm_fold", c_int32),
        ("pool", c_int32), ("reserved0", c_int32),
        ("in2", c_void_p), ("w2", c_void_p), ("cin2", c_int32), ("stride2", c_int32), ("h2", c_int32), ("w2_", c_int32),
        ("in2_stride", c_int64),
    ]


class MbconvDesc(ctypes.Structure):
    """struct af_mbconv_desc"""
    _fields_ = [
        ("in_", c_void_p), ("w1", c_void_p), ("bias1", c_void_p), ("dw_w", c_void_p), ("bias2", c_void_p),
        ("w2", c_void_p), ("bias3", c_void_p), ("residual", c_void_p), ("out", c_void_p),
        ("n", c_int32), ("h", c_int32), ("w_", c_int32), ("cin", c_int32), ("cexp", c_int32), ("cout", c_int32),
        ("stride", c_int32), ("res_stride", c_int64), ("bias1_in_w1", c_int32),
    ]


class MbconvRowsDesc(ctypes.Structure):
    """struct af_mbconv_rows_desc"""
    _fields_ = [
        ("in_", c_void_p), ("w1", c_void_p), ("dwp", c_void_p), ("w2", c_void_p), ("bias3", c_void_p),
        ("residual", c_void_p), ("out", c_void_p),
        ("n", c_int32), ("h", c_int32), ("w_", c_int32), ("cin", c_int32), ("cexp", c_int32), ("cout", c_int32),
        ("stride", c_int32), ("res_stride", c_int64),
        ("in_pix_stride", c_int64), ("in_row_stride", c_int64), ("in_img_stride", c_int64),
    ]


# name -> (restype, argtypes); every symbol declared in include/adafocus_b200.h
SIGNATURES = {
    "af_version": (c_int, []),
    "af_last_error": (c_char_p, []),
    "af_ctx_create": (c_int, [POINTER(c_void_p), c_int]),
    "af_ctx_destroy": (c_int, [c_void_p]),
    "af_ctx_sm_count": (c_int, [c_void_p]),
    "af_plan_begin": (c_int, [c_void_p]),
    "af_plan_end": (c_int, [c_void_p, POINTER(c_void_p)]),
    "af_plan_run": (c_int, [c_void_p, c_void_p]),
    "af_plan_num_launches": (c_int, [c_void_p]),
    "af_plan_destroy": (c_int, [c_void_p]),
    "af_plan_bind_forward": (c_int, [c_void_p, c_void_p, ctypes.c_size_t, c_void_p, ctypes.c_size_t, c_void_p, c_int, c_int,
                                     c_int, c_int, ctypes.c_size_t]),
    "af_workspace_bytes": (ctypes.c_size_t, [c_void_p]),
    "af_gfv_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "af_plan_mark": (c_int, [c_void_p, POINTER(c_int)]),
    "af_plan_mark_elapsed_ms": (c_int, [c_void_p, c_int, c_int, POINTER(c_float)]),
    "af_crop_nchw_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "af_action_to_yx": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "af_stem_im2col": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_int, c_void_p]),
    "af_stem_s2d": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_void_p]),
    "af_stem_conv_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "af_stem_conv3x3s2_c32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "af_conv2d_nhwc_f16": (c_int, [c_void_p, POINTER(ConvDesc), c_void_p]),
    "af_conv_tsm_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "af_mbconv_fused_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "af_mbconv_fused": (c_int, [c_void_p, POINTER(MbconvDesc), c_void_p]),
    "af_mbconv_rows_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "af_mbconv_rows_layout": (c_int, [c_int, c_int, POINTER(c_int32), c_void_p, c_void_p]),
    "af_mbconv_rows": (c_int, [c_void_p, POINTER(MbconvRowsDesc), c_void_p]),
    "af_debug_mbconv_rows_prof": (c_int, [c_void_p]),
    "af_dwconv3x3_nhwc_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_void_p]),
    "af_maxpool3x3s2_nhwc_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "af_avgpool_nhwc_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                    c_void_p]),
    "af_nhwc_f16_to_nchw_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "af_nchw_f32_to_nhwc_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "af_gru_gates": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                             c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "af_split3_f16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "af_gru_sequence": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "af_gru_sequence_tc_supported": (c_int, [c_void_p, c_int, c_int, c_int]),
    "af_gru_sequence_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "af_policy_head": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "af_policy_head_continuous": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                          c_void_p]),
    "af_tsm_shift_nhwc_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "af_tsm_shift_nchw_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "af_consensus_avg": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "af_topk_hits": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "af_softmax_rows": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "af_class_ap": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "af_fill_f32": (c_int, [c_void_p, c_void_p, c_float, c_int64, c_void_p]),
    "af_frames_u8_to_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float),
                                    c_void_p]),
    "af_f32_to_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "af_resize_crop_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def load():
    """dlopen the library and bind every symbol; raises AfError if it is missing (there is no other path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AfError(
            f"{LIB_PATH} not found: build it with `python -m adafocus_b200.build` (nvcc, sm_100a). "
            "adafocus_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("AF_LIB_ALLOW_MISSING") and not hasattr(lib, name):
            continue              # A/B timing against an OLDER build of the ABI (tools only)
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().af_last_error()
        raise AfError(f"{what} failed ({code}): {msg.decode() if msg else ''}")


class Context:
    """Owns one af_ctx (one per device)."""

    def __init__(self, device=0):
        self.lib = load()
        h = c_void_p()
        check(self.lib.af_ctx_create(byref(h), int(device)), "af_ctx_create")
        self.handle = h
        self.device = int(device)
        self.sm_count = self.lib.af_ctx_sm_count(h)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.af_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Plan:
    """A recorded launch sequence (af_plan)."""

    def __init__(self, lib, handle, keepalive):
        self.lib = lib
        self.handle = handle
        self.keepalive = keepalive   # tensors whose pointers the plan replays on
        self.num_launches = lib.af_plan_num_launches(handle)

    def run(self, stream):
        check(self.lib.af_plan_run(self.handle, c_void_p(stream)), "af_plan_run")

    def elapsed_ms(self, mark_a, mark_b):
        """Device time between two marks of the last completed replay (synchronise first)."""
        ms = c_float()
        check(self.lib.af_plan_mark_elapsed_ms(self.handle, int(mark_a), int(mark_b), byref(ms)),
              "af_plan_mark_elapsed_ms")
        return ms.value

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.af_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
