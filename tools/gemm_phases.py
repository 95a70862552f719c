"""Phase timestamps (ns, %globaltimer) of CTA 0 of one GEMM launch: where the fixed per-launch cost goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import ops
dev = torch.device("cuda:0")
buf = torch.zeros(16, dtype=torch.int64, device=dev)
for (m, n, k) in [(1536, 768, 64), (1536, 768, 768), (4096, 768, 768), (12576, 768, 768)]:
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, bias=bias, out=out, block_n=128 | 2048)
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", ops.ptr(buf))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.gemm(a, w, bias=bias, out=out, block_n=128 | 2048); e.record()
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", None)
    t = buf.cpu().tolist()
    names = ["start", "setup done", "first operands", "accumulator done", "epilogue math done", "stores read", "exit"]
    print("%dx%dx%d  events %.1f us: " % (m, n, k, s.elapsed_time(e) * 1e3) + "  ".join("%s +%.2f" % (names[i], (t[i] - t[0]) / 1e3) for i in range(1, 7)))

# epilogue cost of the FFN projection variants on a 128 x 256 tile (single-CTA kernel carries the stamps)
m, n, k = 12576, 3072, 768
a = torch.randn(m, k, device=dev).to(torch.bfloat16)
w = torch.randn(n, k, device=dev).to(torch.bfloat16)
bias = torch.randn(n, device=dev)
out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
pre = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
for label, kw, fl in (("plain", {}, 0),):
    for _ in range(3):
        ops.gemm(a, w, bias=bias, out=out, block_n=256 | 2048 | fl, **kw)
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", ops.ptr(buf))
    ops.gemm(a, w, bias=bias, out=out, block_n=256 | 2048 | fl, **kw)
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", None)
    t = buf.cpu().tolist()
    print("%-28s first tile: accumulator done +%.2f us, epilogue takes %.2f us; CTA total %.2f us" % (
        label, (t[3] - t[0]) / 1e3, (t[4] - t[3]) / 1e3, (t[6] - t[0]) / 1e3))
    nm = ["block1 start", "tmem ld done", "math done (store entry)", "buffer-free barrier", "staged", "proxy fence", "barrier", "bulk store issued"]
    print("      block 1 (+us since accumulator done): " + "  ".join("%s %.2f" % (nm[i], (t[8 + i] - t[3]) / 1e3) for i in range(8)))
for label, kw in (("plain", {}), ("gelu", dict(act="gelu")), ("preact", dict(preact=pre)), ("gelu+preact", dict(act="gelu", preact=pre)),
                  ("gelu' x aux", dict(mul_aux=pre, mul_mode="gelu_grad")), ("residual", dict(residual=pre))):
    for _ in range(2):
        ops.gemm(a, w, bias=bias, out=out, block_n=256 | 2048, **kw)
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", ops.ptr(buf))
    ops.gemm(a, w, bias=bias, out=out, block_n=256 | 2048, **kw)
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", None)
    t = buf.cpu().tolist()
    print("%-12s first tile: accumulator done +%.2f us, epilogue takes %.2f us; CTA total %.2f us" % (
        label, (t[3] - t[0]) / 1e3, (t[4] - t[3]) / 1e3, (t[6] - t[0]) / 1e3))
