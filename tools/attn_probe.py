"""Attention forward / backward timings at the path's shapes (CUDA events, L2 flushed). Not the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import ops
from perf_probe import timeit, dev

B = 32
seed = torch.tensor([7], dtype=torch.int32, device=dev)
for (lq, lk, fr) in [(393, 393, None), (128, 393, None), (48, 393, None), (128, 128, (8, 8)), (40, 40, None),
                     (40, 393, None), (393, 40, None), (177, 40, None), (40, 177, None)]:
    q = torch.randn(B * lq, 2304, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * lk, 2304, device=dev).to(torch.bfloat16)
    d_o = torch.randn(B * lq, 768, device=dev).to(torch.bfloat16)
    dq = torch.empty(B * lq, 768, device=dev, dtype=torch.bfloat16)
    dk = torch.empty(B * lk, 768, device=dev, dtype=torch.bfloat16)
    dv = torch.empty_like(dk)
    fl = 4 * B * 12 * lq * lk * 64
    line = "attn b=%d %4d x %4d:" % (B, lq, lk)
    for drop in (None, (0.1, 3, seed)):
        out, lse = ops.attention_fwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], B, lq, lk, frames=fr, save_lse=True, drop=drop)
        tf = timeit(lambda: ops.attention_fwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], B, lq, lk, frames=fr, save_lse=True, drop=drop))
        tb = timeit(lambda: ops.attention_bwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], out, d_o, lse, dq, dk, dv, B, lq, lk, frames=fr, drop=drop))
        line += "  %s fwd %6.1f us (%5.1f T/s)  bwd %6.1f us (%5.1f T/s)" % ("drop" if drop else "plain", tf * 1e3, fl / tf / 1e9, tb * 1e3, 2.5 * fl / tb / 1e9)
    print(line)
