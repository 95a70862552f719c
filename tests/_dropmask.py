"""Host restatement of the counter-based dropout mask of shg_vqa_b200/csrc/shg_common.cuh (mix32 / drop_key /
drop_pair) so that the GPU tests can evaluate a plain torch reference under the SAME Bernoulli mask the kernels
draw.  Test infrastructure only."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = x.astype(np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & M32
    x ^= x >> np.uint64(16)
    return x


def drop_key(seed, salt):
    v = np.array([(int(seed) ^ ((int(salt) * 0x9E3779B9) & 0xFFFFFFFF)) & 0xFFFFFFFF], dtype=np.uint64)
    return _mix32(v)[0]


def drop_thresh(p):
    return int(np.float32(p) * np.float32(65536.0) + np.float32(0.5))


def keep_from_pairs(key, pair, lane, p):
    """pair: uint64 array of pair indices (already reduced mod 2^32); lane: 0/1 array."""
    h = _mix32((np.uint64(key) + (pair.astype(np.uint64) & M32) * np.uint64(0x85EBCA6B)) & M32)
    lo, hi = h & np.uint64(0xFFFF), h >> np.uint64(16)
    th = np.uint64(drop_thresh(p))
    return np.where(lane == 0, lo >= th, hi >= th)


def keep_flat(seed, salt, rows, n, p):
    """Mask of a row-major [rows, n] tensor (GEMM epilogue / LayerNorm / embedding sites): element idx = row*n+col."""
    idx = np.arange(rows * n, dtype=np.uint64)
    keep = keep_from_pairs(drop_key(seed, salt), (idx >> np.uint64(1)) & M32, (idx & np.uint64(1)).astype(np.int64), p)
    return keep.reshape(rows, n)


def keep_attn(seed, salt, batch, heads, lq, lk, p):
    """Mask of the attention probabilities [batch, heads, lq, lk] (attn.cu: attn_row_pair0)."""
    lk32 = (lk + 31) // 32 * 32
    row = np.arange(batch * heads * lq, dtype=np.uint64)            # ((b*heads + h)*lq + q)
    row_pair0 = ((row & M32) * np.uint64(lk32 >> 1)) & M32
    k = np.arange(lk, dtype=np.uint64)
    pair = (row_pair0[:, None] + (k >> np.uint64(1))[None, :]) & M32
    lane = np.broadcast_to((k & np.uint64(1)).astype(np.int64)[None, :], pair.shape)
    keep = keep_from_pairs(drop_key(seed, salt), pair, lane, p)
    return keep.reshape(batch, heads, lq, lk)
