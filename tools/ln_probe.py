"""GPU time of the LayerNorm kernels at the step's shapes (graph replay of 16 launches).  Not the bench."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import ops
from gemm_probe import graph_time

dev = torch.device("cuda:0")
for rows, n in ((12576, 768), (5632, 768), (1280, 768), (5632, 1536)):
    x = torch.randn(rows, n, device=dev).to(torch.bfloat16)
    dy = torch.randn(rows, n, device=dev).to(torch.bfloat16)
    g = torch.randn(n, device=dev)
    b = torch.randn(n, device=dev)
    y, _, m, r = ops.layernorm_fwd(x, g, b, 1e-12, save_stats=True)
    dg, db, dxs = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    tf = graph_time(lambda i: ops.layernorm_fwd(x, g, b, 1e-12, save_stats=True)) * 1e3
    tb = graph_time(lambda i: ops.layernorm_bwd(dy, x, g, m, r, dg, db, dxsum=dxs)) * 1e3
    print("%6d x %4d  fwd %6.1f us (%5.0f GB/s)   bwd %6.1f us (%5.0f GB/s)" % (
        rows, n, tf, 2 * rows * n * 2 / tf / 1e3, tb, 3 * rows * n * 2 / tb / 1e3), flush=True)
