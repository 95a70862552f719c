"""ctypes binding of libshgvqa.so (the C ABI declared in include/shgvqa.h).

There is no fallback: if the shared library is missing or the device is not sm_100
every entry point raises.  Host code hands raw device pointers and the current
torch CUDA stream to the library; torch only owns memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHG_LIB_PATH") or os.path.join(_HERE, "libshgvqa.so")      # override: A/B builds

c_int, c_i64, c_f32, c_vp = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class ShgError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [("a", c_vp), ("lda", c_i64), ("a_mn_major", c_int),
                ("b", c_vp), ("ldb", c_i64), ("b_mn_major", c_int),
                ("d", c_vp), ("ldd", c_i64),
                ("m", c_int), ("n", c_int), ("k", c_int),
                ("bias", c_vp), ("residual", c_vp), ("ldr", c_i64),
                ("preact", c_vp), ("act", c_int),
                ("mul_aux", c_vp), ("mul_mode", c_int), ("mul_scale", c_f32),
                ("dropout_p", c_f32), ("dropout_salt", ctypes.c_uint32), ("dropout_seed", c_vp),
                ("flags", c_int), ("splits", c_int), ("block_n", c_int),
                ("b2", c_vp), ("bias2", c_vp), ("split_m", c_int), ("d2", c_vp), ("k_split", c_int),
                ("workspace", c_vp), ("workspace_bytes", c_i64), ("max_ctas", c_int)]


class AttnArgs(ctypes.Structure):
    _fields_ = [("q", c_vp), ("ldq", c_i64), ("k", c_vp), ("ldk", c_i64), ("v", c_vp), ("ldv", c_i64),
                ("o", c_vp), ("ldo", c_i64), ("lse", c_vp),
                ("batch", c_int), ("heads", c_int), ("lq", c_int), ("lk", c_int),
                ("scale", c_f32), ("mask_mode", c_int), ("key_mask", c_vp),
                ("q_per_frame", c_int), ("k_per_frame", c_int), ("o_f32", c_vp), ("ldof", c_i64),
                ("dropout_p", c_f32), ("dropout_salt", ctypes.c_uint32), ("dropout_seed", c_vp),
                ("dropout_bits", c_vp)]


class AttnBwdArgs(ctypes.Structure):
    _fields_ = [("q", c_vp), ("ldq", c_i64), ("k", c_vp), ("ldk", c_i64), ("v", c_vp), ("ldv", c_i64),
                ("o", c_vp), ("ldo", c_i64), ("d_o", c_vp), ("lddo", c_i64), ("lse", c_vp), ("delta", c_vp),
                ("dq", c_vp), ("lddq", c_i64), ("dk", c_vp), ("lddk", c_i64), ("dv", c_vp), ("lddv", c_i64),
                ("batch", c_int), ("heads", c_int), ("lq", c_int), ("lk", c_int),
                ("scale", c_f32), ("mask_mode", c_int), ("key_mask", c_vp),
                ("q_per_frame", c_int), ("k_per_frame", c_int), ("o_f32", c_vp), ("ldof", c_i64),
                ("dropout_p", c_f32), ("dropout_salt", ctypes.c_uint32), ("dropout_seed", c_vp),
                ("dropout_bits", c_vp)]


class Conv3dArgs(ctypes.Structure):
    _fields_ = [("x", c_vp), ("w", c_vp), ("bias", c_vp), ("y", c_vp), ("ldy", c_i64),
                ("y_rows_per_clip", c_int), ("y_row_off", c_int), ("preact", c_vp), ("pos_emb", c_vp),
                ("batch", c_int), ("t_in", c_int), ("c_in", c_int), ("c_out", c_int),
                ("dropout_p", c_f32), ("dropout_salt", ctypes.c_uint32), ("dropout_seed", c_vp)]


# name -> argtypes (all return int).  Must list every symbol include/shgvqa.h declares.
SIGNATURES = {
    "shg_version": [],
    "shg_device_check": [c_int],
    "shg_set_attn_bwd_merge": [c_int],
    "shg_set_attn_bwd_chunked": [c_int],
    "shg_set_attn_fwd_chunked_min_lk": [c_int],
    "shg_lsap_batched": [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "shg_match_from_logits": [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                              c_vp, c_vp, c_vp, c_vp, c_vp],
    "shg_set_loss_fwd": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_int, c_vp, c_vp],
    "shg_set_loss_bwd": [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_f32, c_vp, c_f32, c_vp, c_int, c_int, c_vp],
    "shg_gemm": [ctypes.POINTER(GemmArgs), c_vp],
    "shg_debug_gemm_stamps": [c_vp],
    "shg_attn_fwd": [ctypes.POINTER(AttnArgs), c_vp],
    "shg_attn_bwd": [ctypes.POINTER(AttnBwdArgs), c_vp],
    "shg_layernorm_fwd": [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_vp],
    "shg_layernorm_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_f32,
                          ctypes.c_uint32, c_f32, ctypes.c_uint32, c_vp, c_vp],
    "shg_layernorm_fwd_grouped": [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_int,
                                  c_int, c_vp, c_vp, c_vp],
    "shg_layernorm_bwd_grouped": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp,
                                  c_f32, ctypes.c_uint32, c_f32, ctypes.c_uint32, c_vp, c_int, c_vp, c_vp, c_vp, c_vp,
                                  c_vp],
    "shg_colsum_accum": [c_vp, c_i64, c_int, c_int, c_vp, c_vp],
    "shg_embed_ln_fwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_int, c_int,
                         c_int, c_f32, ctypes.c_uint32, c_vp, c_vp],
    "shg_embed_scatter_bwd": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "shg_add_rows": [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_vp],
    "shg_act_bwd": [c_vp, c_vp, c_vp, c_i64, c_int, c_vp],
    "shg_bce_logits": [c_vp, c_i64, c_vp, c_int, c_int, c_f32, c_vp, c_vp, c_i64, c_vp],
    "shg_copy3d": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_i64, c_i64, c_i64, c_i64, c_int, c_vp],
    "shg_grad_sumsq": [c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_vp, c_vp],
    "shg_bert_adam_step": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_int, c_i64, c_vp, c_f32, c_f32, c_vp,
                           c_f32, c_f32, c_f32, c_f32, c_f32, c_int, c_vp],
    "shg_f32_to_bf16": [c_vp, c_vp, c_i64, c_vp],
    "shg_conv3d_fwd": [ctypes.POINTER(Conv3dArgs), c_vp],
    "shg_conv3d_dgrad": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "shg_conv3d_wgrad": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "shg_ncdhw_to_ndhwc": [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    "shg_ndhwc_to_ncdhw": [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    "shg_frontend_cls": [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_f32, ctypes.c_uint32, c_vp, c_vp],
    "shg_frontend_bwd": [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_f32, ctypes.c_uint32, c_vp,
                         c_vp],
}

_lib = None
_checked_devices = set()


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ShgError("libshgvqa.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "— there is no CPU or PyTorch fallback for the hot path" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.shg_last_error.restype = ctypes.c_char_p
        _lib.shg_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = c_int
    return _lib


def ptr(t):
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_device(t):
    """All product entry points go through here: CUDA tensor on an sm_100 device or raise."""
    if not t.is_cuda:
        raise ShgError("shg_vqa_b200 ops need CUDA tensors on a B200 (got %s); there is no CPU path" % t.device)
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _checked_devices:
        check(lib().shg_device_check(idx))
        _checked_devices.add(idx)


def check(rc):
    if rc != 0:
        raise ShgError("libshgvqa error %d: %s" % (rc, lib().shg_last_error().decode()))


# kernels launched per entry point (for the bench's gpu_launches count)
LAUNCHES = {"shg_attn_bwd": 2, "shg_version": 0, "shg_device_check": 0, "shg_set_attn_bwd_merge": 0, "shg_set_attn_bwd_chunked": 0, "shg_set_attn_fwd_chunked_min_lk": 0}
launch_count = 0
profiler = None     # set to a Profiler to time every call with CUDA events on the launching stream


class Profiler:
    """Book-keeping of every entry-point call while installed as `_lib.profiler`: name, roofline kind, algorithmic
    work, shape tag and number of kernels the call launches (bench.py joins these, in order, with CUPTI kernel
    records).  With time_events=True each call is also bracketed by CUDA events on the launching stream — those
    times include the host's enqueue gaps and are only used by the exploratory tools/, never by bench.py."""

    def __init__(self, time_events=False):
        self.time_events = time_events
        self.calls = []        # [name, kind, work, tag, (start_event, end_event) | None, n_kernels]

    def work_by_entry(self):
        agg = {}
        for name, kind, work, tag, _, n_k in self.calls:
            key = "shg_conv3d" if name.startswith("shg_conv3d") else name
            a = agg.setdefault(key, dict(kind=kind, calls=0, work=0.0))
            a["calls"] += 1
            a["work"] += work
        return agg

    def _timed(self, key_fn):
        torch.cuda.synchronize()
        agg = {}
        for name, kind, work, tag, ev, _ in self.calls:
            a = agg.setdefault(key_fn(name, tag), dict(kind=kind, calls=0, ms=0.0, work=0.0))
            a["calls"] += 1
            a["ms"] += ev[0].elapsed_time(ev[1]) if ev is not None else 0.0
            a["work"] += work
        return agg

    def by_tag(self):
        return self._timed(lambda name, tag: (name, tag))

    def summary(self):
        return self._timed(lambda name, tag: name)


def call(name, *args, work=None):
    """work = ("tensor", flops[, tag]) | ("hbm", bytes[, tag]): algorithmic work of this launch, for the roofline."""
    global launch_count
    n_k = LAUNCHES.get(name, 1)
    launch_count += n_k
    if profiler is None:
        check(getattr(lib(), name)(*args))
        return
    ev = None
    if profiler.time_events:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    check(getattr(lib(), name)(*args))
    if ev is not None:
        ev[1].record()
    kind, amount = (work[0], work[1]) if work is not None else ("hbm", 0.0)
    tag = work[2] if work is not None and len(work) > 2 else ""
    profiler.calls.append([name, kind, float(amount), tag, ev, n_k])
