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
elif what == "ln":                         # LayerNorm forward + backward at the relation-layer size (12576 x 768)
    rows, n = B * 393, 768
    x = torch.randn(rows, n, device=dev).to(torch.bfloat16)
    dy = torch.randn(rows, n, device=dev).to(torch.bfloat16)
    g, b_ = torch.ones(n, device=dev), torch.zeros(n, device=dev)
    dg, db, dxs = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for _ in range(3):
        y, _, m, r = ops.layernorm_fwd(x, g, b_, 1e-12, save_stats=True)
        ops.layernorm_bwd(dy, x, g, m, r, dg, db, dxsum=dxs, out_drop=DROP)
elif what == "colsum":                     # bias gradient of the fused QKV projection: column sums of [12576, 2304]
    x = torch.randn(B * 393, 2304, device=dev).to(torch.bfloat16)
    out = torch.zeros(2304, device=dev)
    for _ in range(3):
        ops.colsum_accum(x, out)
elif what == "match":                      # fused matcher on bs32 relation logits (512 frames x 8 queries x 457 classes)
    import bench
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import csr_from_padded
    raw = make_inputs(B, seed=1234, with_feat=False)
    lab, off = [t.to(dev) for t in csr_from_padded(raw["rel_tokens"], raw["rel_lengths"])]
    lg = torch.randn(B * 128, 460, device=dev)[:, :457]
    for _ in range(3):
        m = ops.match_from_logits(lg, lab, off, 8)
        ops.set_loss_fwd(lg, m, 0.1)
elif what == "adam":                       # fused clip + BertAdam over the 191 M live parameters
    from shg_vqa_b200.trainer import HotPathTrainer
    tr = HotPathTrainer(dev, t_total=1000)
    tr.init_weights()
    tr.store.grad.normal_()
    for _ in range(3):
        tr.optimizer_step()
torch.cuda.synchronize()
print("done")
