"""Tensor-level wrappers over the C ABI (no autograd here; see autograd.py).

Every function takes CUDA tensors, launches on the current torch stream and
returns torch tensors it allocated; none falls back to PyTorch math.
"""
import torch

from . import _lib
from ._lib import AttnArgs, AttnBwdArgs, Conv3dArgs, GemmArgs, call, ptr, require_device, stream

ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2, "tanh": 3}
MUL = {None: 0, "gelu_grad": 1, "relu_mask": 2}
GEMM_OUT_F32 = 1
GEMM_ACCUMULATE = 2
BF16 = torch.bfloat16
import os as _os
ATTN_BWD_MERGE = _os.environ.get("SHG_ATTN_MERGE", "1") != "0"     # mirrors the library's default (see shg_set_attn_bwd_merge)
_PAIR_FLAG = {"0": 2048, "1": 1024}.get(_os.environ.get("SHG_GEMM_PAIR", ""), 0)   # debugging: forbid / force CTA-pair tiles


# Scratch of the GEMM's tail-split schedule (include/shgvqa.h: shg_gemm_args.workspace).  A buffer may serve one launch
# at a time, so there is one per CONCURRENCY LANE: the engine runs its launches on a main stream and three side streams
# (language side, action side, weight gradients); launches of one lane are stream-ordered, launches of different lanes
# overlap.  Lanes rather than stream ids, because the main lane is whatever stream is current (default stream, warm-up
# stream, graph-capture stream).  `workspace_lane(i)` selects the lane for the calls made inside it.
_GEMM_WS = {}
_gemm_lane = 0
GEMM_WS_BYTES = 1024 + 10 * 2 * 128 * 256 * 4         # up to 10 split tiles of 256 x 256 (the bs32 step needs 8)
GEMM_TAIL_SPLIT = _os.environ.get("SHG_GEMM_TAIL_SPLIT", "1") != "0"


class workspace_lane:
    def __init__(self, lane):
        self.lane = lane

    def __enter__(self):
        global _gemm_lane
        self.prev, _gemm_lane = _gemm_lane, self.lane

    def __exit__(self, *a):
        global _gemm_lane
        _gemm_lane = self.prev


# Cap on the persistent GEMM grid (shg_gemm_args.max_ctas) for the calls made inside `gemm_cta_cap(n)`: the
# data-parallel step leaves the SMs its in-flight all-reduce owns to NCCL (trainer.HotPathTrainer.train_step).
_gemm_max_ctas = 0


class gemm_cta_cap:
    def __init__(self, n):
        self.n = int(n)

    def __enter__(self):
        global _gemm_max_ctas
        self.prev, _gemm_max_ctas = _gemm_max_ctas, self.n

    def __exit__(self, *a):
        global _gemm_max_ctas
        _gemm_max_ctas = self.prev


def _gemm_workspace(dev):
    if not GEMM_TAIL_SPLIT:
        return None
    key = (dev.index, _gemm_lane)
    ws = _GEMM_WS.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            return None                     # first seen inside a graph capture: plain schedule for this launch
        ws = torch.zeros(GEMM_WS_BYTES, device=dev, dtype=torch.uint8)
        _GEMM_WS[key] = ws
    return ws


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "row-major 2-D tensor (any leading dimension) expected"
    return t.stride(0)


def gemm(a, b, *, a_t=False, b_t=False, bias=None, act=None, residual=None, preact=None, mul_aux=None,
         mul_mode=None, mul_scale=1.0, drop=None, out=None, out_f32=False, accumulate=False, splits=0, block_n=0,
         b2=None, bias2=None, split_m=0, d2=None, k_split=0):
    # drop = (p, salt, seed_tensor) enables training-mode dropout on the output (before the residual add)
    """D = epilogue(A . B^T).  a: [M,K] (or stored [K,M] with a_t), b: [N,K] (or stored [K,N] with b_t); bf16."""
    require_device(a)
    assert a.dtype == BF16 and b.dtype == BF16
    if a_t:
        k, m = a.shape
    else:
        m, k = a.shape
    if b_t:
        kb, n = b.shape
    else:
        n, kb = b.shape
    assert k == kb, (a.shape, b.shape, a_t, b_t)
    if out is None:
        out = torch.empty(m, n, device=a.device, dtype=torch.float32 if (out_f32 or accumulate) else BF16)
    assert out.shape == (m, n)
    args = GemmArgs()
    args.a, args.lda, args.a_mn_major = ptr(a), _ld(a), int(a_t)
    args.b, args.ldb, args.b_mn_major = ptr(b), _ld(b), int(b_t)
    args.d, args.ldd = ptr(out), _ld(out)
    args.m, args.n, args.k = m, n, k
    args.bias = ptr(bias)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n and bias.is_contiguous()
    args.residual = ptr(residual)
    args.ldr = _ld(residual) if residual is not None else 0
    if preact is not None:
        assert preact.dtype == BF16 and _ld(preact) == _ld(out) and preact.shape == out.shape
    args.preact = ptr(preact)
    args.act = ACT[act]
    if mul_aux is not None:
        assert mul_aux.dtype == BF16 and _ld(mul_aux) == _ld(out) and mul_aux.shape == out.shape
    args.mul_aux, args.mul_mode, args.mul_scale = ptr(mul_aux), MUL[mul_mode], mul_scale
    if drop is not None and drop[0] > 0:
        args.dropout_p, args.dropout_salt, args.dropout_seed = drop[0], drop[1], ptr(drop[2])
    if b2 is not None:
        # grouped product: rows >= split_m of A go through the second weight / bias (same shapes)
        assert b2.dtype == BF16 and b2.shape == b.shape and _ld(b2) == _ld(b) and 0 < split_m < m
        assert (bias2 is None) == (bias is None)
        args.b2, args.bias2, args.split_m = ptr(b2), ptr(bias2), split_m
    if d2 is not None:
        # grouped split-K accumulation: K rows >= k_split add into d2 (second layer's weight gradient)
        assert accumulate and d2.dtype == torch.float32 and d2.shape == out.shape and _ld(d2) == _ld(out)
        assert 0 < k_split < k and splits == 0
        args.d2, args.k_split = ptr(d2), k_split
    ws = _gemm_workspace(a.device) if (m > 256 * 74 // 2 and not accumulate) else None
    if ws is not None:
        args.workspace, args.workspace_bytes = ptr(ws), ws.numel()
    args.max_ctas = _gemm_max_ctas
    args.flags = (GEMM_OUT_F32 if out.dtype == torch.float32 else 0) | (GEMM_ACCUMULATE if accumulate else 0)
    args.splits, args.block_n = splits, (block_n | _PAIR_FLAG if block_n == 0 and _PAIR_FLAG else block_n)
    call("shg_gemm", args, stream(), work=("tensor", 2.0 * m * n * k, "%dx%dx%d %s%s%s" % (m, n, k, "T" if a_t else "N", "T" if b_t else "N", " acc" if accumulate else "")))
    return out


def _set_drop(a, drop):
    if drop is not None and drop[0] > 0:
        a.dropout_p, a.dropout_salt, a.dropout_seed = drop[0], drop[1], ptr(drop[2])


def attention_fwd(q, k, v, batch, lq, lk, heads=12, scale=0.125, key_mask=None, frames=None, out=None,
                  save_lse=False, drop=None):
    """q [batch*lq, >=heads*64], k/v [batch*lk, ...] bf16 views (row stride = ld, unit column stride).
    key_mask: fp32 [batch, lk] additive.  frames=(q_per_frame, k_per_frame) selects the block-causal mask."""
    require_device(q)
    if out is None:
        out = torch.empty(batch * lq, heads * 64, device=q.device, dtype=BF16)
    lse = torch.empty(batch, heads, lq, device=q.device, dtype=torch.float32) if save_lse else None
    o32 = torch.empty(batch * lq, heads * 64, device=q.device, dtype=torch.float32) if save_lse else None
    a = AttnArgs()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = ptr(q), _ld(q), ptr(k), _ld(k), ptr(v), _ld(v)
    a.o, a.ldo, a.lse = ptr(out), _ld(out), ptr(lse)
    a.batch, a.heads, a.lq, a.lk, a.scale = batch, heads, lq, lk, scale
    a.mask_mode = 1 if key_mask is not None else (2 if frames is not None else 0)
    if key_mask is not None:
        assert key_mask.dtype == torch.float32 and key_mask.is_contiguous() and key_mask.shape == (batch, lk)
    a.key_mask = ptr(key_mask)
    a.q_per_frame, a.k_per_frame = frames if frames is not None else (0, 0)
    a.o_f32, a.ldof = ptr(o32), heads * 64
    _set_drop(a, drop)
    bits = None
    if save_lse and drop is not None and drop[0] > 0:
        # keep mask as bit words for the backward (1 bit per probability: 2 % of the bf16 score matrix it stands for)
        bits = torch.empty(batch, heads, lq, (lk + 31) // 32, device=q.device, dtype=torch.int32)
        a.dropout_bits = ptr(bits)
    call("shg_attn_fwd", a, stream(), work=("tensor", 4.0 * batch * heads * lq * lk * 64, "%dx%d" % (lq, lk)))
    return out, (lse, o32, bits) if save_lse else None


def attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, batch, lq, lk, heads=12, scale=0.125, key_mask=None,
                  frames=None, drop=None):
    """Writes dq/dk/dv (bf16 views with the same head layout as q/k/v).  `lse` is what attention_fwd returned
    with save_lse: (log-sum-exp, fp32 O); a bare lse tensor selects the two-sweep delta."""
    require_device(q)
    o32 = bits = None
    if isinstance(lse, tuple):
        lse, o32, bits = lse
    delta = torch.empty_like(lse)
    a = AttnBwdArgs()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = ptr(q), _ld(q), ptr(k), _ld(k), ptr(v), _ld(v)
    a.o, a.ldo, a.d_o, a.lddo = ptr(o), _ld(o), ptr(d_o), _ld(d_o)
    a.lse, a.delta = ptr(lse), ptr(delta)
    a.dq, a.lddq, a.dk, a.lddk, a.dv, a.lddv = ptr(dq), _ld(dq), ptr(dk), _ld(dk), ptr(dv), _ld(dv)
    a.batch, a.heads, a.lq, a.lk, a.scale = batch, heads, lq, lk, scale
    a.mask_mode = 1 if key_mask is not None else (2 if frames is not None else 0)
    a.key_mask = ptr(key_mask)
    a.q_per_frame, a.k_per_frame = frames if frames is not None else (0, 0)
    a.o_f32, a.ldof = ptr(o32), (o32.stride(0) if o32 is not None else 0)
    _set_drop(a, drop)
    if drop is not None and drop[0] > 0:
        a.dropout_bits = ptr(bits)
    call("shg_attn_bwd", a, stream(), work=("tensor", 10.0 * batch * heads * lq * lk * 64, "%dx%d" % (lq, lk)))
    if lq <= 128 and o32 is not None and ATTN_BWD_MERGE:
        _lib.launch_count -= 1          # single-query-tile problems take one merged launch instead of two
        if _lib.profiler is not None:
            _lib.profiler.calls[-1][5] = 1


def layernorm_fwd(x, gamma, beta, eps, residual=None, addend=None, addend_rows=0, save_stats=False, group=None):
    """group = (split_row, gamma2, beta2): rows >= split_row are normalised with the second parameter set."""
    require_device(x)
    rows, n = x.shape
    assert x.dtype == BF16 and x.is_contiguous() and gamma.dtype == torch.float32
    y = torch.empty_like(x)
    y2 = torch.empty_like(x) if addend is not None else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty_like(mean) if save_stats else None
    split, g2, b2 = group if group is not None else (0x7fffffff, None, None)
    call("shg_layernorm_fwd_grouped", ptr(x), ptr(residual), ptr(gamma), ptr(beta), eps, ptr(y), ptr(y2), ptr(addend),
         addend_rows, ptr(mean), ptr(rstd), rows, n, split, ptr(g2), ptr(b2), stream(),
         work=("hbm", 2.0 * rows * n * (2 + (residual is not None) + 2 * (addend is not None)), "%dx%d" % (rows, n)))
    return y, y2, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dy2=None, residual=None, dxsum=None, out_drop=None,
                  in_drop=None, group=None):
    """out_drop / in_drop = (p, salt, seed_tensor).  With out_drop returns (dx, dx_drop).
    group = (split_row, gamma2, dgamma2, dbeta2, dxsum2): second parameter / gradient set for rows >= split_row."""
    require_device(dy)
    rows, n = x.shape
    dx = torch.empty_like(x)
    od = out_drop if (out_drop is not None and out_drop[0] > 0) else None
    idr = in_drop if (in_drop is not None and in_drop[0] > 0) else None
    dxd = torch.empty_like(x) if od is not None else None
    seed = od[2] if od is not None else (idr[2] if idr is not None else None)
    split, g2, dg2, db2, dxs2 = group if group is not None else (0x7fffffff, None, None, None, None)
    call("shg_layernorm_bwd_grouped", ptr(dy), ptr(dy2), ptr(x), ptr(residual), ptr(gamma), ptr(mean), ptr(rstd),
         ptr(dx), ptr(dgamma), ptr(dbeta), ptr(dxsum), rows, n, ptr(dxd), od[0] if od else 0.0, od[1] if od else 0,
         idr[0] if idr else 0.0, idr[1] if idr else 0, ptr(seed), split, ptr(g2), ptr(dg2), ptr(db2), ptr(dxs2),
         stream(),
         work=("hbm", 2.0 * rows * n * (3 + (dy2 is not None) + (residual is not None)), "%dx%d" % (rows, n)))
    return (dx, dxd) if od is not None else dx


def colsum_accum(x, out):
    require_device(x)
    rows, n = x.shape
    call("shg_colsum_accum", ptr(x), _ld(x), rows, n, ptr(out), stream(), work=("hbm", 2.0 * rows * n, "%dx%d" % (rows, n)))


def add_rows(x, add):
    require_device(x)
    rows, n = x.shape
    y = torch.empty_like(x)
    call("shg_add_rows", ptr(x), ptr(add), add.shape[0], ptr(y), rows, n, stream())
    return y


def act_bwd(dy, aux, act):
    require_device(dy)
    dx = torch.empty_like(dy)
    call("shg_act_bwd", ptr(dy), ptr(aux), ptr(dx), dy.numel(), ACT[act], stream(), work=("hbm", 6.0 * dy.numel()))
    return dx


def bce_logits(x, target, out, grad_scale=1.0, want_grad=True):
    """x fp32 [rows, cols] (row stride = ld). Adds the loss into out[0]; returns d loss/dx as bf16 (ld % 8 == 0)."""
    require_device(x)
    rows, cols = x.shape
    dx = None
    ld = (cols + 7) // 8 * 8
    if want_grad:
        dx = torch.zeros(rows, ld, device=x.device, dtype=BF16)[:, :cols]
    call("shg_bce_logits", ptr(x), _ld(x), ptr(target), rows, cols, grad_scale, ptr(out), ptr(dx), ld, stream())
    return dx


def copy3d(src, dst, n0, n1, chunk, s0, s1, d0, d1, add=None, accumulate=False):
    """Strided block copy over raw bf16 storage; strides in elements (see include/shgvqa.h)."""
    require_device(src)
    call("shg_copy3d", ptr(src), ptr(dst), ptr(add), n0, n1, chunk, s0, s1, d0, d1, int(accumulate), stream(),
         work=("hbm", 2.0 * n0 * n1 * chunk * (2 + bool(accumulate))))


def to_bf16(x, out=None):
    require_device(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    call("shg_f32_to_bf16", ptr(x), ptr(out), x.numel(), stream(), work=("hbm", 6.0 * x.numel()))
    return out


# ----------------------------------------------------------------------------- Conv3d visual front-end

def pack_conv_weight(w):
    """Reference layout [c_out, c_in, 5, 3, 3] -> packed [c_out, 45*c_in] (tap-major, channel-minor)."""
    return w.permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1).contiguous()


def unpack_conv_weight(wp, c_in):
    """Packed [c_out, 45*c_in] -> a VIEW with the reference's shape [c_out, c_in, 5, 3, 3]."""
    return wp.view(wp.shape[0], 5, 3, 3, c_in).permute(0, 4, 1, 2, 3)


def ncdhw_to_ndhwc(x, out=None):
    """fp32 [B, C, T, 7, 7] (backbone output) -> bf16 channels-last [B, T, 7, 7, C]."""
    require_device(x)
    b, c, t, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and h == 7 and w == 7
    if out is None:
        out = torch.empty(b, t, h, w, c, device=x.device, dtype=BF16)
    call("shg_ncdhw_to_ndhwc", ptr(x), ptr(out), b, c, t * h * w, stream(), work=("hbm", 6.0 * x.numel()))
    return out


def ndhwc_to_ncdhw(x, out=None):
    """bf16 [B, T, 7, 7, C] -> fp32 [B, C, T, 7, 7] (gradient handed back to the backbone)."""
    require_device(x)
    b, t, h, w, c = x.shape
    assert x.dtype == BF16 and x.is_contiguous()
    if out is None:
        out = torch.empty(b, c, t, h, w, device=x.device, dtype=torch.float32)
    call("shg_ndhwc_to_ncdhw", ptr(x), ptr(out), b, c, t * h * w, stream(), work=("hbm", 6.0 * x.numel()))
    return out


def conv3d_fwd(x, wp, bias, *, out=None, rows_per_clip=0, row_off=0, preact=None, pos_emb=None, drop=None):
    """GELU(Conv3d(ZeroPad2d(1)(x))) with kernel (5,3,3).  x bf16 [B, T, 7, 7, c_in]; wp bf16 packed [c_out, 45*c_in].
    out: bf16 2-D token matrix the result rows are scattered into (see include/shgvqa.h) or None for a fresh
    [B*(T-4)*49, c_out]; preact: True to save the pre-activation (returned), or a tensor to fill."""
    require_device(x)
    b, t, h, w, c_in = x.shape
    c_out = wp.shape[0]
    assert x.dtype == BF16 and wp.dtype == BF16 and x.is_contiguous() and wp.is_contiguous()
    assert wp.shape[1] == 45 * c_in and h == 7 and w == 7
    t_out = t - 4
    if out is None:
        out = torch.empty(b * t_out * 49, c_out, device=x.device, dtype=BF16)
    if preact is True:
        preact = torch.empty(b * t_out * 49, c_out, device=x.device, dtype=BF16)
    a = Conv3dArgs()
    a.x, a.w, a.bias, a.y, a.ldy = ptr(x), ptr(wp), ptr(bias), ptr(out), _ld(out)
    a.y_rows_per_clip, a.y_row_off = rows_per_clip, row_off
    a.preact, a.pos_emb = ptr(preact), ptr(pos_emb)
    a.batch, a.t_in, a.c_in, a.c_out = b, t, c_in, c_out
    _set_drop(a, drop)
    call("shg_conv3d_fwd", a, stream(),
         work=("tensor", 2.0 * b * t_out * 49 * c_out * 45 * c_in, "conv fwd %d->%d T%d" % (c_in, c_out, t)))
    return out, preact


def conv3d_dgrad(dy, wp, c_in, preact=None, out=None):
    """dy bf16 [B, T_out, 7, 7, c_out] -> dx bf16 [B, T_out+4, 7, 7, c_in] (x gelu'(preact) if given)."""
    require_device(dy)
    b, t_out, h, w, c_out = dy.shape
    assert dy.dtype == BF16 and dy.is_contiguous() and wp.shape == (c_out, 45 * c_in)
    t_in = t_out + 4
    if out is None:
        out = torch.empty(b, t_in, 7, 7, c_in, device=dy.device, dtype=BF16)
    call("shg_conv3d_dgrad", ptr(dy), ptr(wp), ptr(preact), ptr(out), b, t_in, c_in, c_out, stream(),
         work=("tensor", 2.0 * b * t_out * 49 * c_in * 45 * c_out, "conv dgrad %d->%d T%d" % (c_in, c_out, t_in)))
    return out


def conv3d_wgrad(dy, x, dw):
    """dw (fp32 packed [c_out, 45*c_in]) += grad of the conv weight; dy bf16 [B, T-4, 7, 7, c_out], x bf16 [B, T, 7, 7, c_in]."""
    require_device(dy)
    b, t, h, w, c_in = x.shape
    c_out = dy.shape[-1]
    assert dy.dtype == BF16 and x.dtype == BF16 and dy.is_contiguous() and x.is_contiguous()
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == c_out * 45 * c_in
    assert dy.shape[:4] == (b, t - 4, 7, 7)
    call("shg_conv3d_wgrad", ptr(dy), ptr(x), ptr(dw), b, t, c_in, c_out, stream(),
         work=("tensor", 2.0 * b * (t - 4) * 49 * c_out * 45 * c_in, "conv wgrad %d->%d T%d" % (c_in, c_out, t)))
    return dw


def frontend_cls(cls, pe, tokens, batch, rows_per_clip, drop=None):
    require_device(tokens)
    n = tokens.shape[1]
    d = drop if (drop is not None and drop[0] > 0) else None
    call("shg_frontend_cls", ptr(cls), ptr(pe), ptr(tokens), _ld(tokens), batch, rows_per_clip, n,
         d[0] if d else 0.0, d[1] if d else 0, ptr(d[2]) if d else None, stream())


def frontend_bwd(d_tok, preact, d_pe, d_cls, batch, rows_per_clip, drop=None):
    """Returns d_pre bf16 [batch*(rows_per_clip-1), n]; accumulates into d_pe [rows_per_clip, n] and d_cls [n]."""
    require_device(d_tok)
    n = d_tok.shape[1]
    d_pre = torch.empty(batch * (rows_per_clip - 1), n, device=d_tok.device, dtype=BF16)
    d = drop if (drop is not None and drop[0] > 0) else None
    call("shg_frontend_bwd", ptr(d_tok), _ld(d_tok), ptr(preact), ptr(d_pre), ptr(d_pe), ptr(d_cls), batch,
         rows_per_clip, n, d[0] if d else 0.0, d[1] if d else 0, ptr(d[2]) if d else None, stream(),
         work=("hbm", 2.0 * batch * rows_per_clip * n * 3))
    return d_pre


# ----------------------------------------------------------------------------- matching

def lsap_batched(cost, col_off, q_f):
    """cost fp32 [n_frames, q_f, ld] device; col_off int32 [n_frames+1] device."""
    require_device(cost)
    n_frames = cost.shape[0]
    assert cost.dtype == torch.float32 and cost.is_contiguous() and cost.shape[1] == q_f
    dev = cost.device
    rows = torch.empty(n_frames, q_f, device=dev, dtype=torch.int32)
    cols = torch.empty_like(rows)
    n_match = torch.empty(n_frames, device=dev, dtype=torch.int32)
    status = torch.empty_like(n_match)
    call("shg_lsap_batched", ptr(cost), n_frames, q_f, cost.shape[2], ptr(col_off), ptr(rows), ptr(cols),
         ptr(n_match), ptr(status), stream())
    return rows, cols, n_match, status


def match_from_logits(logits, labels, label_off, q_f, want_targets=True):
    """logits [n_frames*q_f, n_cls] fp32|bf16 (row stride = ld); labels int32 CSR."""
    require_device(logits)
    n_rows, n_cls = logits.shape
    n_frames = n_rows // q_f
    dev = logits.device
    rows = torch.empty(n_frames, q_f, device=dev, dtype=torch.int32)
    cols = torch.empty_like(rows)
    n_match = torch.empty(n_frames, device=dev, dtype=torch.int32)
    status = torch.empty_like(n_match)
    row_max = torch.empty(n_rows, device=dev, dtype=torch.float32)
    row_lse = torch.empty_like(row_max)
    row_arg = torch.empty(n_rows, device=dev, dtype=torch.int32)
    tgt = torch.empty(n_rows, device=dev, dtype=torch.int32) if want_targets else None
    call("shg_match_from_logits", ptr(logits), int(logits.dtype == BF16), n_frames, q_f, n_cls, _ld(logits),
         ptr(labels), ptr(label_off), ptr(rows), ptr(cols), ptr(n_match), ptr(status), ptr(row_max), ptr(row_lse),
         ptr(row_arg), ptr(tgt), stream(), work=("hbm", float(logits.element_size()) * n_rows * n_cls))
    return dict(rows=rows, cols=cols, n_match=n_match, status=status, row_max=row_max, row_lse=row_lse,
                row_argmax=row_arg, target_classes=tgt)


def set_loss_fwd(logits, m, eos_coef, out=None, err_out=None):
    """err_out: fp32 device scalar that counts frames the matcher refused (status != 0) and out-of-range target
    classes — the inputs the reference raises on (see include/shgvqa.h)."""
    n_rows, n_cls = logits.shape
    if out is None:
        out = torch.zeros(4, device=logits.device, dtype=torch.float32)
    status = m.get("status") if err_out is not None else None
    call("shg_set_loss_fwd", ptr(logits), int(logits.dtype == BF16), n_rows, n_cls, _ld(logits),
         ptr(m["target_classes"]), ptr(m["row_lse"]), ptr(m["row_argmax"]), eos_coef, ptr(out), ptr(status),
         status.numel() if status is not None else 0, ptr(err_out), stream())
    return out


def set_loss_bwd(logits, m, eos_coef, sums, grad_scale=1.0, out_dtype=BF16):
    n_rows, n_cls = logits.shape
    ld = (n_cls + 7) // 8 * 8
    buf = torch.zeros(n_rows, ld, device=logits.device, dtype=out_dtype)
    d = buf[:, :n_cls]
    call("shg_set_loss_bwd", ptr(logits), int(logits.dtype == BF16), n_rows, n_cls, _ld(logits),
         ptr(m["target_classes"]), ptr(m["row_lse"]), eos_coef, ptr(sums), grad_scale, ptr(d),
         int(out_dtype == BF16), ld, stream())
    return d
