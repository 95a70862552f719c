"""Conv3d front-end kernels vs torch fp32 on the same bf16-rounded operands, plus timings (GPU box only).

    python tools/conv_probe.py [batch]
"""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from shg_vqa_b200 import ops

BF16 = torch.bfloat16


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    for (c_in, c_out, t_in) in ((768, 768, 12), (2048, 768, 16)):
        x = torch.randn(B, c_in, t_in, 7, 7, device=dev, generator=g).relu_()
        w = torch.randn(c_out, c_in, 5, 3, 3, device=dev, generator=g) * 0.01
        bias = torch.randn(c_out, device=dev, generator=g) * 0.1
        xb = x.to(BF16).float()
        wb = w.to(BF16).float()
        xc = ops.ncdhw_to_ndhwc(x)
        assert torch.equal(xc, x.permute(0, 2, 3, 4, 1).to(BF16)), "layout kernel"
        back = ops.ndhwc_to_ncdhw(xc)
        assert torch.equal(back, xc.permute(0, 4, 1, 2, 3).float()), "layout kernel (back)"
        wp = ops.pack_conv_weight(w).to(BF16)
        y, pre = ops.conv3d_fwd(xc, wp, bias, preact=True)
        torch.cuda.synchronize()
        if B <= 4:
            xr = xb.clone().requires_grad_(True)
            wr = wb.clone().requires_grad_(True)
            pre_ref = F.conv3d(F.pad(xr, (1, 1, 1, 1)), wr, bias)
            y_ref = F.gelu(pre_ref)
            t_out = t_in - 4
            pre_ref_cl = pre_ref.permute(0, 2, 3, 4, 1).reshape(-1, c_out)
            y_ref_cl = y_ref.permute(0, 2, 3, 4, 1).reshape(-1, c_out)
            print("conv %d->%d T%d B%d: fwd preact rel %.2e  gelu rel %.2e" % (c_in, c_out, t_in, B, rel(pre, pre_ref_cl),
                                                                            rel(y, y_ref_cl)))
            dy = torch.randn(B, t_out, 7, 7, c_out, device=dev, generator=g).to(BF16)
            # backward through the conv only (dy is the gradient wrt the pre-activation)
            pre_ref.backward(dy.float().permute(0, 4, 1, 2, 3))
            dx = ops.conv3d_dgrad(dy, wp, c_in)
            dw = torch.zeros(c_out, 45 * c_in, device=dev)
            ops.conv3d_wgrad(dy, xc, dw)
            torch.cuda.synchronize()
            print("   dgrad rel %.2e   wgrad rel %.2e" % (rel(dx, xr.grad.permute(0, 2, 3, 4, 1)),
                                                         rel(ops.unpack_conv_weight(dw, c_in), wr.grad)))
        else:
            t_out = t_in - 4
            dy = torch.randn(B, t_out, 7, 7, c_out, device=dev, generator=g).to(BF16)
            dw = torch.zeros(c_out, 45 * c_in, device=dev)
        fl = 2.0 * B * t_out * 49 * c_out * 45 * c_in
        ms = timeit(lambda: ops.conv3d_fwd(xc, wp, bias, out=y, preact=pre))
        print("   fwd   %.3f ms  %.0f TFLOP/s" % (ms, fl / ms / 1e9))
        ms = timeit(lambda: ops.conv3d_dgrad(dy, wp, c_in))
        print("   dgrad %.3f ms  %.0f TFLOP/s" % (ms, fl / ms / 1e9))
        ms = timeit(lambda: ops.conv3d_wgrad(dy, xc, dw))
        print("   wgrad %.3f ms  %.0f TFLOP/s" % (ms, fl / ms / 1e9))
        if B >= 8:
            xh = x.to(BF16).contiguous(memory_format=torch.channels_last_3d)
            wh = w.to(BF16).contiguous(memory_format=torch.channels_last_3d)
            ms = timeit(lambda: F.conv3d(F.pad(xh, (1, 1, 1, 1)), wh, bias.to(BF16)))
            print("   cuDNN bf16 channels_last_3d fwd %.3f ms  %.0f TFLOP/s" % (ms, fl / ms / 1e9))


if __name__ == "__main__":
    main()
