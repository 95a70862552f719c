import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import ops
dev = torch.device("cuda:0")
buf = torch.zeros(16, dtype=torch.int64, device=dev)
for (m, n, k, bn) in [(5632, 768, 768, 256), (5632, 768, 768, 128), (5632, 2048, 768, 256), (5632, 768, 2048, 256)]:
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, bias=bias, out=out, block_n=bn | 2048)
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", ops.ptr(buf))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ops.gemm(a, w, bias=bias, out=out, block_n=bn | 2048); e.record()
    torch.cuda.synchronize()
    ops.call("shg_debug_gemm_stamps", None)
    t = buf.cpu().tolist()
    names = ["start", "setup done", "first operands", "accumulator done", "epilogue math done", "stores read", "exit"]
    print("%dx%dx%d bn%d events %.1f us: " % (m, n, k, bn, s.elapsed_time(e) * 1e3) + "  ".join("%s +%.2f" % (names[i], (t[i] - t[0]) / 1e3) for i in range(1, 7)))
    en = ["acc visible", "block entry", "first tmem ld", "block staged", "proxy fence", "wg barrier"]
    print("      epilogue of warpgroup 0 (+us since start): " + "  ".join("%s %.2f" % (en[i], (t[8 + i] - t[0]) / 1e3) for i in range(6)))
