"""Conv3d visual front-end (lxrt/modeling_capsbert.py:989-996, 1037-1073) on the CUDA path.

Kernel level: the implicit-GEMM forward / dgrad / wgrad, the layout and token-assembly kernels against torch fp32
on the SAME bf16-rounded operands (tolerance: bf16 output rounding, 4e-3 rel-L2; wgrad accumulates and stores fp32:
1e-5), dropout under the regenerated mask.  Model level: the engine run from the seeded backbone features against
the golden vectors of the UNMODIFIED reference (which ran its own Conv3d front-end): logits, losses, and the
gradient norms of the front-end's parameters; and element-wise against torch autograd through the oracle's
front-end fed the engine's own upstream gradient."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _dropmask import keep_flat

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16
VFC = "lxrt_encoder.model.bert.encoder.visn_fc"


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("c_in,c_out,t_in,batch", [(768, 768, 12, 2), (2048, 768, 16, 2), (256, 256, 8, 3)])
def test_conv3d_kernels_vs_torch(c_in, c_out, t_in, batch):
    from shg_vqa_b200 import ops
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(c_in + t_in)
    x = torch.randn(batch, c_in, t_in, 7, 7, device=dev, generator=g).relu_()
    w = torch.randn(c_out, c_in, 5, 3, 3, device=dev, generator=g) * 0.01
    bias = torch.randn(c_out, device=dev, generator=g) * 0.1
    xc = ops.ncdhw_to_ndhwc(x)
    assert torch.equal(xc, x.permute(0, 2, 3, 4, 1).to(BF16))                 # layout kernels are exact
    assert torch.equal(ops.ndhwc_to_ncdhw(xc), xc.permute(0, 4, 1, 2, 3).float())
    wp = ops.pack_conv_weight(w).to(BF16)
    assert torch.equal(ops.unpack_conv_weight(wp, c_in), w.to(BF16))
    y, pre = ops.conv3d_fwd(xc, wp, bias, preact=True)
    xr = xc.permute(0, 4, 1, 2, 3).float().requires_grad_(True)
    wr = ops.unpack_conv_weight(wp, c_in).float().contiguous().requires_grad_(True)
    pre_ref = F.conv3d(F.pad(xr, (1, 1, 1, 1)), wr, bias)
    y_ref = F.gelu(pre_ref)
    cl = lambda t: t.permute(0, 2, 3, 4, 1).reshape(-1, t.shape[1])
    assert rel_l2(pre, cl(pre_ref)) <= 4e-3
    assert rel_l2(y, cl(y_ref)) <= 4e-3
    t_out = t_in - 4
    dy = torch.randn(batch, t_out, 7, 7, c_out, device=dev, generator=g).to(BF16)
    pre_ref.backward(dy.float().permute(0, 4, 1, 2, 3))
    dx = ops.conv3d_dgrad(dy, wp, c_in)
    assert rel_l2(dx, xr.grad.permute(0, 2, 3, 4, 1)) <= 4e-3
    dw = torch.full((c_out, 45 * c_in), 0.5, device=dev)                      # accumulates into what is there
    ops.conv3d_wgrad(dy, xc, dw)
    assert rel_l2(ops.unpack_conv_weight(dw - 0.5, c_in), wr.grad) <= 1e-5
    # dgrad fused with the GELU derivative of the layer below
    aux = torch.randn(batch * t_in * 49, c_in, device=dev, generator=g).to(BF16)
    dx2 = ops.conv3d_dgrad(dy, wp, c_in, preact=aux)
    a = aux.float().requires_grad_(True)
    F.gelu(a).sum().backward()
    ref = xr.grad.permute(0, 2, 3, 4, 1).reshape(-1, c_in) * a.grad
    assert rel_l2(dx2.view(-1, c_in), ref) <= 4e-3


def test_conv3d_rejects_unsupported_geometry():
    from shg_vqa_b200 import ops
    from shg_vqa_b200._lib import ShgError
    dev = torch.device("cuda:0")
    x = torch.zeros(1, 10, 7, 7, 256, device=dev, dtype=BF16)                 # t_in not a multiple of 4
    w = torch.zeros(256, 45 * 256, device=dev, dtype=BF16)
    with pytest.raises(ShgError):
        ops.conv3d_fwd(x, w, None)
    with pytest.raises(ShgError):
        ops.conv3d_fwd(torch.zeros(1, 8, 7, 7, 128, device=dev, dtype=BF16),
                       torch.zeros(256, 45 * 128, device=dev, dtype=BF16), None)


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_token_assembly_fwd_bwd(p):
    """Second conv's epilogue + cls kernel = dropout(cat(cls, gelu(conv)) + pe); frontend_bwd is its backward."""
    from shg_vqa_b200 import ops
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device=dev).manual_seed(7)
    B, T, C = 3, 12, 256
    nv = 1 + (T - 4) * 49
    x = torch.randn(B, T, 7, 7, C, device=dev, generator=g).to(BF16)
    wp = (torch.randn(C, 45 * C, device=dev, generator=g) * 0.02).to(BF16)
    bias = torch.randn(C, device=dev, generator=g) * 0.1
    pe = torch.randn(nv + 3, C, device=dev, generator=g)
    cls = torch.randn(C, device=dev, generator=g)
    seed = torch.tensor([1234], device=dev, dtype=torch.int32)
    drop = (p, 17, seed) if p > 0 else None
    tokens = torch.empty(B * nv, C, device=dev, dtype=BF16)
    _, pre = ops.conv3d_fwd(x, wp, bias, out=tokens, rows_per_clip=nv, row_off=1, preact=True, pos_emb=pe, drop=drop)
    ops.frontend_cls(cls, pe, tokens, B, nv, drop=drop)
    keep = torch.from_numpy(keep_flat(1234, 17, B * nv, C, p)).to(dev) if p > 0 else torch.ones(B * nv, C, device=dev, dtype=torch.bool)
    pre_r = pre.float().view(B, nv - 1, C).requires_grad_(True)
    pe_r, cls_r = pe.clone().requires_grad_(True), cls.clone().requires_grad_(True)
    tok_r = torch.cat([cls_r.expand(B, 1, C), F.gelu(pre_r)], dim=1) + pe_r[:nv]
    tok_r = (tok_r.reshape(B * nv, C) * keep) / (1.0 - p)
    assert rel_l2(tokens, tok_r) <= 4e-3
    kept = float(keep.float().mean())
    assert abs(kept - (1.0 - p)) < 0.01
    assert bool(((tokens == 0) | keep).all())                                  # dropped elements are exactly zero
    d_tok = torch.randn(B * nv, C, device=dev, generator=g).to(BF16)
    tok_r.backward(d_tok.float())
    d_pe = torch.zeros(nv, C, device=dev)
    d_cls = torch.zeros(C, device=dev)
    d_pre = ops.frontend_bwd(d_tok, pre, d_pe, d_cls, B, nv, drop=drop)
    assert rel_l2(d_pre, pre_r.grad) <= 4e-3
    assert rel_l2(d_pe, pe_r.grad[:nv]) <= 1e-5
    assert rel_l2(d_cls, cls_r.grad) <= 1e-5


@pytest.fixture(scope="module")
def run(oracle_weights, inputs_b2, golden_hgqa):
    from shg_vqa_b200.trainer import HotPathTrainer, csr_targets
    dev = torch.device("cuda:0")
    tr = HotPathTrainer(dev, frontend=True, backbone_grad=True)
    tr.load_state_dict(oracle_weights, strict=True)
    inp = {k: inputs_b2[k].to(dev) for k in ("feat", "input_ids", "input_mask", "segment_ids", "rel_segment_ids",
                                             "act_segment_ids")}
    batch = {"target": inputs_b2["target"].to(dev)}
    batch["rel_labels"], batch["rel_off"] = csr_targets(inputs_b2["rel_labels"], dev)
    batch["act_labels"], batch["act_off"] = csr_targets(inputs_b2["act_labels"], dev)
    out, ctx = tr.engine.forward(inp, train=True)
    # the reference's assignment held fixed (see tests/test_gpu_model.py)
    for fam, q_f, labels in (("rel", 8, inputs_b2["rel_labels"]), ("act", 3, inputs_b2["act_labels"])):
        rows, cols = golden_hgqa[fam + "_rows"], golden_hgqa[fam + "_cols"]
        tgt = torch.zeros(rows.shape[0], q_f, dtype=torch.int32)
        for f in range(rows.shape[0]):
            k = int((rows[f] >= 0).sum())
            tgt[f, torch.from_numpy(rows[f, :k])] = labels[f][torch.from_numpy(cols[f, :k])].to(torch.int32)
        batch[fam + "_target_classes"] = tgt.reshape(-1).to(dev)
    scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, batch)
    d_visn = tr.engine.backward(ctx, d_hg, d_rel, d_act, need_dvisn=True)
    torch.cuda.synchronize()
    return tr, out, scal.clone(), ctx, d_visn


def test_model_from_backbone_features_vs_reference(run, golden_hgqa, oracle_weights, inputs_b2):
    from oracle import shg_oracle as so
    tr, out, scal, ctx, _ = run
    g = golden_hgqa
    # tokens vs the oracle's fp32 front-end (itself pinned to the reference through the golden logits)
    tok_ref = so.visual_front_end(oracle_weights, VFC, inputs_b2["feat"]).reshape(-1, 768)
    x0, pre1, a1, pre2, dv, nv = ctx["frontend"]
    assert nv == 393
    visn_r = out["visn_r"].float().cpu().view(2, 393, 768)
    assert rel_l2(visn_r[:, ::4], torch.from_numpy(g["visn_r"])) <= 2e-2
    for name in ("logit", "rel_logits", "act_logits", "hg_logit"):
        mine = out[name].float().cpu().reshape(g[name].shape)
        ref = torch.from_numpy(g[name])
        assert float((mine - ref).abs().max()) <= 3e-2 * float(ref.abs().max()), name
    losses = tr.losses_from_scalars(scal)
    assert abs(losses["total"] - float(g["loss_total"])) <= 1e-2 * abs(float(g["loss_total"]))
    del tok_ref


def test_frontend_tokens_vs_oracle(oracle_weights, inputs_b2):
    from oracle import shg_oracle as so
    from shg_vqa_b200.trainer import HotPathTrainer
    dev = torch.device("cuda:0")
    tr = HotPathTrainer(dev, frontend=True)
    tr.load_state_dict(oracle_weights, strict=True)
    tok, _ = tr.engine.frontend_fwd(inputs_b2["feat"].to(dev), save=False)
    tok_ref = so.visual_front_end(oracle_weights, VFC, inputs_b2["feat"]).reshape(-1, 768)
    assert rel_l2(tok.float().cpu(), tok_ref) <= 6e-3          # two stacked bf16 convolutions, K = 92160 / 34560


def test_frontend_gradients_vs_reference(run, golden_hgqa):
    """Gradient norms of the front-end's own parameters against the unmodified reference's (5 %, the tolerance of
    every other tensor on the path), nothing where the reference has no gradient."""
    tr, _, _, _, _ = run
    g = golden_hgqa
    st = tr.store
    seen = 0
    for n, ref in zip((str(x) for x in g["grad_names"]), g["grad_norms"]):
        if ".visn_fc." not in n:
            continue
        mine = float(st.g(n).double().norm())
        assert abs(mine - ref) <= 5e-2 * max(ref, 1e-6), (n, mine, float(ref))
        seen += 1
    assert seen == 6                                           # conv.1/.4 weight+bias, cls_token, pe.weight
    for n in g["nograd_names"]:
        if ".visn_fc." in str(n):
            assert float(st.g(str(n)).abs().max()) == 0.0, n


def test_frontend_backward_elementwise_vs_autograd(run, oracle_weights, inputs_b2):
    """Same upstream gradient (the engine's own d_visn) through torch-fp32 autograd of the oracle's front-end:
    every front-end gradient element-wise, and the gradient handed to the backbone."""
    from oracle import shg_oracle as so
    tr, _, _, ctx, d_visn = run
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    keys = [VFC + s for s in (".conv.1.weight", ".conv.1.bias", ".conv.4.weight", ".conv.4.bias", ".cls_token",
                              ".position_encoding.pe.weight")]
    sd = {k: oracle_weights[k].to(dev).clone().requires_grad_(True) for k in keys}
    feat = inputs_b2["feat"].to(dev).clone().requires_grad_(True)
    tok = so.visual_front_end(sd, VFC, feat).reshape(-1, 768)
    tok.backward(d_visn.float())
    st = tr.store
    for k in keys:
        tol = 2e-2 if k.endswith("weight") and "conv" in k else 1e-2
        assert rel_l2(st.g(k), sd[k].grad) <= tol, (k, rel_l2(st.g(k), sd[k].grad))
    assert ctx["d_feat"].shape == feat.shape and ctx["d_feat"].dtype == torch.float32
    assert rel_l2(ctx["d_feat"], feat.grad) <= 2e-2


def test_state_dict_keeps_reference_conv_layout(run, oracle_weights):
    tr, _, _, _, _ = run
    sd = tr.state_dict()
    for k in (VFC + ".conv.1.weight", VFC + ".conv.4.weight"):
        assert tuple(sd[k].shape) == tuple(oracle_weights[k].shape)
        assert torch.equal(sd[k].cpu(), oracle_weights[k])
        # what the kernels read is the packed permutation of the same storage
        assert torch.equal(tr.store.packed(tr.store.master, k).cpu(),
                           oracle_weights[k].permute(0, 2, 3, 4, 1).reshape(oracle_weights[k].shape[0], -1))
