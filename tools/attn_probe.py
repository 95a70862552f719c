"""GPU time of the attention kernels at the path's (Lq, Lk) problems, B = 32, dropout on (as the bench step runs
them): 16 back-to-back launches captured in a CUDA graph and replayed (no host launch latency).  Forward: whole-row
kernel vs the chunked two-CTA-per-SM kernel; backward: as dispatched.  Not the bench."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import _lib, ops
from gemm_probe import graph_time

dev = torch.device("cuda:0")
B = 32
seed = torch.tensor([7], dtype=torch.int32, device=dev)
DROP = None if "--nodrop" in sys.argv else (0.1, 3, seed)
SHAPES = [(393, 393), (128, 393), (48, 393), (40, 393), (128, 128), (48, 48), (40, 40), (393, 40), (40, 177), (177, 40)]
for lq, lk in SHAPES:
    q = torch.randn(B * lq, 2304, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * lk, 2304, device=dev).to(torch.bfloat16)
    frames = (8, 8) if (lq, lk) == (128, 128) else ((3, 3) if (lq, lk) == (48, 48) else None)
    res = {}
    for name, min_lk in (("whole-row", 100000), ("chunked", 1)):
        if name == "whole-row" and lk > 416:
            continue
        _lib.check(_lib.lib().shg_set_attn_fwd_chunked_min_lk(min_lk))
        f = lambda i: ops.attention_fwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], B, lq, lk, frames=frames,
                                        save_lse=True, drop=DROP)
        res[name] = graph_time(f) * 1e3
    _lib.check(_lib.lib().shg_set_attn_fwd_chunked_min_lk(257))
    out, lse = ops.attention_fwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], B, lq, lk, frames=frames, save_lse=True,
                                 drop=DROP)
    d_o = torch.randn(B * lq, 768, device=dev).to(torch.bfloat16)
    dq = torch.empty(B * lq, 768, device=dev, dtype=torch.bfloat16)
    dk = torch.empty(B * lk, 768, device=dev, dtype=torch.bfloat16)
    dv = torch.empty(B * lk, 768, device=dev, dtype=torch.bfloat16)
    fb = lambda i: ops.attention_bwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], out, d_o, lse, dq, dk, dv, B, lq, lk,
                                     frames=frames, drop=DROP)
    tb = graph_time(fb) * 1e3
    _lib.check(_lib.lib().shg_set_attn_bwd_chunked(0))
    tb_old = graph_time(fb) * 1e3
    _lib.check(_lib.lib().shg_set_attn_bwd_chunked(1))
    _lib.check(_lib.lib().shg_set_attn_bwd_merge(0))
    tb_nomerge = graph_time(fb) * 1e3
    _lib.check(_lib.lib().shg_set_attn_bwd_merge(1))
    fl = 4.0 * B * 12 * lq * lk * 64
    print("%3d x %3d  fwd whole-row %7.1f us  chunked %7.1f us (%6.1f TF/s)   bwd %7.1f us (%6.1f TF/s)  first-gen bwd %7.1f us  merge off %7.1f us" % (
        lq, lk, res.get("whole-row", float("nan")), res["chunked"], fl / res["chunked"] / 1e6, tb, 2.5 * fl / tb / 1e6, tb_old, tb_nomerge),
        flush=True)
