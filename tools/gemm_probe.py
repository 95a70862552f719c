"""True GPU time per GEMM call at the path's shapes: each case is captured 16x back to back into a CUDA graph
(rotating over distinct operand buffers > L2) and replayed, so Python/ctypes launch latency is out of the picture, as
it is in the captured training step.  cuBLAS (torch.matmul) in the same harness is printed beside it.  Not the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shg_vqa_b200 import ops

dev = torch.device("cuda:0")
REP = 16


def graph_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / REP)
    return sorted(ts)[2]


def case(m, n, k, mode, extra=None):
    """mode: NN fwd (x[m,k] w[n,k]^T), NT dgrad (dy[m,k] w[k,n] stored [k,n]... as b_t), TT wgrad accumulate."""
    nb = max(2, min(8, int(300e6 / (2.0 * (m * k + n * k + m * n)) + 1)))
    kw = dict(extra or {})
    if mode == "NN":
        a = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nb)]
        w = [torch.randn(n, k, device=dev).to(torch.bfloat16) for _ in range(nb)]
        bias = torch.randn(n, device=dev)
        out = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
        if kw.pop("gelu_only", False):
            f = lambda i: ops.gemm(a[i % nb], w[i % nb], bias=bias, act="gelu", out=out[i % nb], **kw)
        elif kw.pop("pre_only", False):
            pre = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
            f = lambda i: ops.gemm(a[i % nb], w[i % nb], bias=bias, preact=pre[i % nb], out=out[i % nb], **kw)
        elif kw.pop("gelu", False):
            pre = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
            f = lambda i: ops.gemm(a[i % nb], w[i % nb], bias=bias, act="gelu", preact=pre[i % nb], out=out[i % nb], **kw)
        elif kw.pop("res", False):
            f = lambda i: ops.gemm(a[i % nb], w[i % nb], bias=bias, residual=out[(i + 1) % nb], out=out[i % nb], **kw)
        else:
            f = lambda i: ops.gemm(a[i % nb], w[i % nb], bias=bias, out=out[i % nb], **kw)
        r = lambda i: torch.matmul(a[i % nb], w[i % nb].t(), out=out[i % nb])
    elif mode == "NT":      # dx[m,n] = dy[m,k] . w[k,n]  with w stored [k, n] row-major  (b_t)
        a = [torch.randn(m, k, device=dev).to(torch.bfloat16) for _ in range(nb)]
        w = [torch.randn(k, n, device=dev).to(torch.bfloat16) for _ in range(nb)]
        out = [torch.empty(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
        f = lambda i: ops.gemm(a[i % nb], w[i % nb], b_t=True, out=out[i % nb], **kw)
        r = lambda i: torch.matmul(a[i % nb], w[i % nb], out=out[i % nb])
    else:                   # TT acc: dw[m,n] += dy[k,m]^T x[k,n]
        a = [torch.randn(k, m, device=dev).to(torch.bfloat16) for _ in range(nb)]
        w = [torch.randn(k, n, device=dev).to(torch.bfloat16) for _ in range(nb)]
        out = [torch.zeros(m, n, device=dev, dtype=torch.float32) for _ in range(nb)]
        outb = [torch.zeros(m, n, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
        f = lambda i: ops.gemm(a[i % nb], w[i % nb], a_t=True, b_t=True, out=out[i % nb], accumulate=True, **kw)
        r = lambda i: torch.matmul(a[i % nb].t(), w[i % nb], out=outb[i % nb])
    t, t2 = graph_time(f), graph_time(r)
    fl = 2.0 * m * n * k
    print("%-3s %6d x %5d x %5d %-22s: %7.1f us %7.1f TF/s | cuBLAS %7.1f us %7.1f TF/s | x%.2f" % (
        mode, m, n, k, str(extra or ""), t * 1e3, fl / t / 1e9, t2 * 1e3, fl / t2 / 1e9, t / t2), flush=True)


if __name__ == "__main__":
    B = 32
    sel = sys.argv[1] if len(sys.argv) > 1 else "all"
    if sel == "ab":     # staged TMA-store epilogue vs direct stores
        for (m, n, k, mode, ex) in [(12576, 3072, 768, "NN", dict(gelu=True)), (12576, 3072, 768, "NN", None),
                                    (12576, 768, 3072, "NN", dict(res=True)), (12576, 768, 768, "NN", dict(res=True)),
                                    (12576, 2304, 768, "NN", None), (4096, 768, 768, "NN", None),
                                    (4096, 2048, 768, "NN", None), (1536, 768, 768, "NN", None),
                                    (1280, 3072, 768, "NN", None), (12576, 768, 2304, "NT", None)]:
            case(m, n, k, mode, ex)
            case(m, n, k, mode, dict(ex or {}, block_n=4096))
    if sel in ("tune", "tune2"):
        shapes2 = [(5632, 768, 768, "NN"), (5632, 1536, 768, "NN"), (5632, 2048, 768, "NN"), (5632, 768, 2048, "NN"),
                   (5632, 768, 768, "NT"), (5632, 768, 1536, "NT"), (5632, 2048, 768, "NT"), (5632, 768, 2048, "NT")]
        shapes = [(12576, 768, 768, "NN"), (12576, 768, 768, "NT"), (12576, 768, 3072, "NN"), (12576, 768, 3072, "NT"),
                  (12576, 768, 2304, "NT"), (12576, 3072, 768, "NT"), (12576, 768, 15360, "NT"),
                  (4096, 768, 768, "NN"), (4096, 768, 768, "NT"), (4096, 1536, 768, "NN"), (4096, 2048, 768, "NN"),
                  (4096, 768, 2048, "NN"), (4096, 2048, 768, "NT"), (4096, 768, 1536, "NT"), (4096, 457, 1536, "NN"),
                  (1536, 768, 768, "NN"), (1536, 1536, 768, "NN"), (1536, 2048, 768, "NN"), (1536, 768, 2048, "NN"),
                  (1280, 2304, 768, "NN"), (1280, 3072, 768, "NN"), (1280, 768, 3072, "NN"), (1280, 768, 768, "NN"),
                  (1280, 3072, 768, "NT"), (5664, 3072, 768, "NN"), (5664, 768, 3072, "NN"), (5664, 2304, 768, "NN"),
                  (5664, 768, 768, "NN"), (5664, 768, 2304, "NT"),
                  (768, 768, 4096, "TT"), (768, 2048, 4096, "TT"), (2048, 768, 4096, "TT"), (768, 768, 1536, "TT"),
                  (768, 3072, 1280, "TT"), (3072, 768, 1280, "TT"), (2304, 768, 1280, "TT"), (768, 768, 1280, "TT"),
                  (768, 3072, 5664, "TT"), (2304, 768, 5664, "TT"), (3072, 768, 12576, "TT"), (768, 3072, 12576, "TT"),
                  (2304, 768, 12576, "TT"), (768, 768, 12576, "TT"), (1536, 768, 12576, "TT"), (1536, 768, 4096, "TT")]
        if sel == "tune2":      # stacked decoder rows (both decoders in lock-step)
            shapes = shapes2
        import io, contextlib, re
        cub = float('nan')
        for (m, n, k, mode) in shapes:
            res = {}
            cfgs = [("auto", 0)] + [("s%d" % bn, bn | 2048) for bn in (64, 128, 192, 256)] + \
                   [("p%d" % bn, bn | 1024) for bn in (128, 192, 256)] + [("p256notail", 256 | 1024 | 16384)]
            for name, flag in cfgs:
                if name == "p192" and mode != "NN":
                    continue
                buf = io.StringIO()
                try:
                    with contextlib.redirect_stdout(buf):
                        case(m, n, k, mode, dict(block_n=flag) if flag else None)
                    mt = re.search(r":\s+([\d.]+) us\s+[\d.]+ TF/s \| cuBLAS\s+([\d.]+) us", buf.getvalue())
                    res[name], cub = float(mt.group(1)), float(mt.group(2))
                except Exception as e:
                    res[name] = float("nan")
            best = min((v, k2) for k2, v in res.items() if v == v and k2 != "auto")
            print("%-2s %6d %5d %6d  auto %6.1f  best %-5s %6.1f  cuBLAS %6.1f | %s" % (
                mode, m, n, k, res["auto"], best[1], best[0], cub,
                " ".join("%s=%.1f" % (k2, v) for k2, v in res.items() if k2 != "auto")), flush=True)
    if sel == "fixed":      # T(K) = a + b K: fixed per-launch cost vs per-k-block cost
        for (m, n) in [(1536, 768), (4096, 768), (12576, 768)]:
            for k in (64, 256, 768, 1536, 3072):
                case(m, n, k, "NN", dict(block_n=128 | 2048))
    if sel == "gelu":
        for ex in (None, dict(gelu_only=True), dict(pre_only=True), dict(gelu=True)):
            case(12576, 3072, 768, "NN", ex)
            case(5664, 3072, 768, "NN", ex)
    if sel in ("all", "big"):
        for (m, n, k, mode, ex) in [
                (12576, 3072, 768, "NN", dict(gelu=True)), (12576, 3072, 768, "NN", None), (12576, 3072, 768, "NT", None),
                (12576, 768, 3072, "NN", dict(res=True)), (12576, 768, 3072, "NT", None), (12576, 2304, 768, "NN", None),
                (12576, 768, 2304, "NT", None), (12576, 768, 768, "NN", dict(res=True)), (12576, 768, 768, "NT", None),
                (12576, 15360, 768, "NN", None), (12576, 768, 15360, "NT", None),
                (3072, 768, 12576, "TT", None), (768, 3072, 12576, "TT", None), (2304, 768, 12576, "TT", None),
                (768, 768, 12576, "TT", None), (1536, 768, 12576, "TT", None)]:
            case(m, n, k, mode, ex)
    if sel in ("all", "mid"):
        for (m, n, k, mode, ex) in [
                (4096, 768, 768, "NN", None), (4096, 768, 768, "NT", None), (768, 768, 4096, "TT", None),
                (4096, 1536, 768, "NN", None), (4096, 2048, 768, "NN", None), (4096, 768, 2048, "NN", None),
                (768, 2048, 4096, "TT", None), (2048, 768, 4096, "TT", None),
                (1536, 768, 768, "NN", None), (1536, 768, 768, "NT", None), (768, 768, 1536, "TT", None),
                (1536, 2048, 768, "NN", None), (1536, 768, 2048, "NN", None), (768, 2048, 1536, "TT", None),
                (1280, 2304, 768, "NN", None), (1280, 3072, 768, "NN", None), (1280, 768, 3072, "NN", None),
                (1280, 768, 768, "NN", None), (768, 3072, 1280, "TT", None), (3072, 768, 1280, "TT", None),
                (768, 768, 1280, "TT", None),
                (5664, 3072, 768, "NN", None), (5664, 768, 3072, "NN", None), (5664, 2304, 768, "NN", None),
                (5664, 768, 768, "NN", None), (768, 3072, 5664, "TT", None),
                (4096, 457, 1536, "NN", None), (32, 768, 1536, "NN", None)]:
            case(m, n, k, mode, ex)
