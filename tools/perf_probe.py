"""Quick per-kernel timing on a B200 (CUDA events, L2-flushed between iterations). Not the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_case(m, n, k, **kw):
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = torch.randn(n, k, device=dev).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm(a, w, bias=bias, out=out, **kw))
    t2 = timeit(lambda: torch.nn.functional.linear(a, w, bias.to(torch.bfloat16)))
    print("gemm %6d x %5d x %5d %-14s: %8.1f us  %7.1f TFLOP/s   (cuBLAS %8.1f us %7.1f TFLOP/s)" % (
        m, n, k, str(kw), t * 1e3, 2 * m * n * k / t / 1e9, t2 * 1e3, 2 * m * n * k / t2 / 1e9))


def attn_case(b, lq, lk, frames=None):
    q = torch.randn(b * lq, 2304, device=dev).to(torch.bfloat16)
    kv = torch.randn(b * lk, 2304, device=dev).to(torch.bfloat16)
    t = timeit(lambda: ops.attention_fwd(q[:, :768], kv[:, 768:1536], kv[:, 1536:], b, lq, lk, frames=frames))
    fl = 4 * b * 12 * lq * lk * 64
    print("attn b=%d %4d x %4d: %8.1f us  %7.1f TFLOP/s" % (b, lq, lk, t * 1e3, fl / t / 1e9))


if __name__ == "__main__":
    B = 32
    for (m, n, k) in [(B * 393, 2304, 768), (B * 393, 768, 768), (B * 393, 3072, 768), (B * 393, 768, 3072),
                      (B * 40, 2304, 768), (B * 40, 3072, 768), (B * 128, 2304, 768), (B * 128, 2048, 768),
                      (B * 128, 768, 2048), (B * 393, 15360, 768), (B * 128, 457, 1536), (8192, 8192, 8192)]:
        gemm_case(m, n, k)
    for (m, n, k) in [(B * 393, 3072, 768), (B * 393, 768, 768), (B * 393, 768, 3072), (B * 393, 2304, 768), (4096, 2048, 768), (8192, 8192, 8192)]:
        for bn in (128 | 2048, 192 | 2048, 256 | 2048, 128 | 1024, 256 | 1024):
            gemm_case(m, n, k, block_n=bn)
    for (lq, lk, fr) in [(393, 393, None), (40, 40, None), (128, 393, None), (48, 393, None), (128, 128, (8, 8)),
                         (393, 40, None), (40, 393, None), (177, 40, None)]:
        attn_case(B, lq, lk, fr)
    x = torch.randn(B * 393, 768, device=dev).to(torch.bfloat16)
    g = torch.ones(768, device=dev); bt = torch.zeros(768, device=dev)
    t = timeit(lambda: ops.layernorm_fwd(x, g, bt, 1e-12))
    print("layernorm %d rows: %.1f us  %.0f GB/s" % (x.shape[0], t * 1e3, 2 * x.numel() * 2 / t / 1e6))
