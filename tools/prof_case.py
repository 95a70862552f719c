"""Tiny single-kernel driver for ncu captures: python tools/prof_case.py attn|gemm|ln"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import ops
dev = torch.device("cuda:0")
B = 32
what = sys.argv[1]
seed = torch.tensor([7], dtype=torch.int32, device=dev)
DROP = (0.1, 3, seed)            # training mode, as in the bench step
if what == "attn":
    lq = int(sys.argv[2]) if len(sys.argv) > 2 else 393
    lk = int(sys.argv[3]) if len(sys.argv) > 3 else 393
    q = torch.randn(B * lq, 2304, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * lk, 2304, device=dev).to(torch.bfloat16)
    for _ in range(3):
        ops.attention_fwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], B, lq, lk, save_lse=True, drop=DROP)
elif what == "attn_bwd":
    lq = int(sys.argv[2]) if len(sys.argv) > 2 else 393
    lk = int(sys.argv[3]) if len(sys.argv) > 3 else 393
    q = torch.randn(B * lq, 2304, device=dev).to(torch.bfloat16)
    kv = torch.randn(B * lk, 2304, device=dev).to(torch.bfloat16)
    d_o = torch.randn(B * lq, 768, device=dev).to(torch.bfloat16)
    dq = torch.empty(B * lq, 768, device=dev, dtype=torch.bfloat16)
    dk = torch.empty(B * lk, 768, device=dev, dtype=torch.bfloat16)
    dv = torch.empty(B * lk, 768, device=dev, dtype=torch.bfloat16)
    out, lse = ops.attention_fwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], B, lq, lk, save_lse=True, drop=DROP)
    for _ in range(3):
        ops.attention_bwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], out, d_o, lse, dq, dk, dv, B, lq, lk, drop=DROP)
elif what == "gemm":
    m, n, k = [int(v) for v in sys.argv[2:5]] if len(sys.argv) > 4 else (B * 393, 3072, 768)
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    act = sys.argv[5] if len(sys.argv) > 5 else None
    pre = torch.empty(m, n, device=dev, dtype=torch.bfloat16) if act == "gelu" else None     # FFN1 as the step runs it
    for _ in range(3):
        ops.gemm(a, w, bias=bias, act=act, preact=pre)
elif what in ("conv_fwd", "conv_wgrad", "conv_dgrad"):
    c_in, c_out, t_in = (2048, 768, 16) if len(sys.argv) < 3 or sys.argv[2] == "1" else (768, 768, 12)
    x = torch.randn(B, t_in, 7, 7, c_in, device=dev).relu_().to(torch.bfloat16)
    wp = (torch.randn(c_out, 45 * c_in, device=dev) * 0.01).to(torch.bfloat16)
    bias = torch.randn(c_out, device=dev)
    dy = torch.randn(B, t_in - 4, 7, 7, c_out, device=dev).to(torch.bfloat16)
    dw = torch.zeros(c_out, 45 * c_in, device=dev)
    for _ in range(3):
        if what == "conv_fwd":
            ops.conv3d_fwd(x, wp, bias, preact=True)
        elif what == "conv_wgrad":
            ops.conv3d_wgrad(dy, x, dw)
        else:
            ops.conv3d_dgrad(dy, wp, c_in)
torch.cuda.synchronize()
print("done")
