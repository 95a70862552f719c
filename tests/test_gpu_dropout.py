"""Training-mode dropout of the hot path (reference: nn.Dropout at lxrt/modeling_capsbert.py:311,339,381,429,483,
nn.MultiheadAttention(dropout=) and dropout1-3 at lxrt/transformer.py:192-204).  The reference draws its masks from
torch's Philox stream, which a different kernel decomposition cannot reproduce; what is checked here is that every
kernel applies ONE well-defined iid Bernoulli(1-p) mask with 1/(1-p) rescaling, identically in forward and backward:
the masks are regenerated on the host (tests/_dropmask.py) and a plain torch fp32 evaluation under that exact mask
is the reference.  Tolerances as in test_gpu_kernels.py (bf16 operands)."""
import numpy as np
import pytest
import torch

from _dropmask import keep_attn, keep_flat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from shg_vqa_b200 import ops as o
    return o


def _dev():
    return torch.device("cuda:0")


def _seed(v):
    return torch.tensor([v], dtype=torch.int32, device=_dev())


@pytest.mark.parametrize("act", [None, "relu"])
def test_gemm_epilogue_dropout(ops, act):
    g = torch.Generator().manual_seed(5)
    m, n, k, p = 1000, 768, 768, 0.15
    a = torch.randn(m, k, generator=g).to(torch.bfloat16).to(_dev())
    w = (torch.randn(n, k, generator=g) * 0.04).to(torch.bfloat16).to(_dev())
    bias = (torch.randn(n, generator=g) * 0.1).to(_dev())
    res = torch.randn(m, n, generator=g).to(torch.bfloat16).to(_dev())
    seed = _seed(99)
    out = ops.gemm(a, w, bias=bias, act=act, residual=res, drop=(p, 3, seed))
    keep = torch.from_numpy(keep_flat(99, 3, m, n, p)).to(_dev())
    z = a.float() @ w.float().t() + bias
    if act == "relu":
        z = torch.relu(z)
    ref = torch.where(keep, z / (1 - p), torch.zeros_like(z)) + res.float()
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    # dropped positions carry exactly the residual
    assert torch.equal(out[~keep], res[~keep])
    # relu-mask dgrad of a dropped+rescaled hidden: d = (h > 0) ? d / (1-p) : 0
    if act == "relu":
        h = ops.gemm(a, w, bias=bias, act="relu", drop=(p, 3, seed))
        dy = torch.randn(m, k, generator=g).to(torch.bfloat16).to(_dev())
        d = ops.gemm(dy, w, mul_aux=h, mul_mode="relu_mask", mul_scale=1 / (1 - p))
        ref = torch.where(h.float() > 0, (dy.float() @ w.float().t()) / (1 - p), torch.zeros(m, n, device=_dev()))
        assert (d.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


def test_layernorm_bwd_dropout_hooks(ops):
    g = torch.Generator().manual_seed(6)
    rows, n, p_out, p_in = 999, 768, 0.1, 0.15
    x = (torch.randn(rows, n, generator=g) * 2).to(torch.bfloat16).to(_dev())
    gamma = (1 + 0.1 * torch.randn(n, generator=g)).to(_dev())
    beta = (0.1 * torch.randn(n, generator=g)).to(_dev())
    y, _, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-12, save_stats=True)
    dy = torch.randn(rows, n, generator=g).to(torch.bfloat16).to(_dev())
    seed = _seed(17)
    z = lambda: torch.zeros(n, device=_dev())
    # plain
    dg0, db0 = z(), z()
    dx0 = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dg0, db0)
    # x = dropout(dense_out) + residual  → second output masked and rescaled, dxsum accumulates it
    dg, db, dxs = z(), z(), z()
    dx, dxd = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dg, db, dxsum=dxs, out_drop=(p_out, 5, seed))
    keep = torch.from_numpy(keep_flat(17, 5, rows, n, p_out)).to(_dev())
    assert torch.equal(dx, dx0)
    ref = torch.where(keep, dx0.float() / (1 - p_out), torch.zeros_like(dx0.float()))
    assert (dxd.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert torch.equal(dxd[~keep], torch.zeros_like(dxd[~keep]))
    assert torch.allclose(dxs, dxd.float().sum(0), rtol=1e-2, atol=0.25)      # fp32 sums vs sums of bf16-rounded values
    # y = dropout(LN(x)) → the incoming gradient is masked first
    dg2, db2 = z(), z()
    dx2 = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dg2, db2, in_drop=(p_in, 9, seed))
    keep_in = torch.from_numpy(keep_flat(17, 9, rows, n, p_in)).to(_dev())
    xs = x.float().requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref_y = torch.nn.functional.layer_norm(xs, (n,), gp, bp, 1e-12)
    ref_y.backward(torch.where(keep_in, dy.float() / (1 - p_in), torch.zeros_like(ref_y)))
    assert (dx2.float() - xs.grad).abs().max().item() <= 1e-2 * xs.grad.abs().max().item()
    assert torch.allclose(dg2, gp.grad, rtol=1e-3, atol=1e-2)
    assert torch.allclose(db2, bp.grad, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("lq,lk,mode,p", [(393, 393, "none", 0.1), (40, 393, "none", 0.1), (393, 40, "key", 0.1),
                                          (128, 128, "causal8", 0.15), (48, 393, "none", 0.15),
                                          (177, 40, "key", 0.1)])
def test_attention_dropout_fwd_bwd(ops, lq, lk, mode, p):
    """o = dropout(softmax(s)) v and its gradients, against torch autograd under the same mask."""
    b, seed_v, salt = 2, 4242, 11
    g = torch.Generator().manual_seed(lq * 977 + lk)
    qkv_q = torch.randn(b * lq, 2304, generator=g).to(torch.bfloat16).to(_dev())
    qkv_k = torch.randn(b * lk, 2304, generator=g).to(torch.bfloat16).to(_dev())
    q, k, v = qkv_q[:, :768], qkv_k[:, 768:1536], qkv_k[:, 1536:]
    key_mask, frames, add = None, None, None
    if mode == "key":
        valid = torch.randint(max(1, lk // 3), lk + 1, (b,), generator=g)
        m01 = (torch.arange(lk)[None, :] < valid[:, None]).float()
        key_mask = ((1 - m01) * -10000.0).to(_dev())
        add = key_mask[:, None, None, :]
    elif mode.startswith("causal"):
        per = int(mode[6:])
        frames = (per, per)
        f = torch.arange(lq) // per
        add = torch.zeros(lq, lk)
        add[f[None, :lk] > f[:, None]] = float("-inf")
        add = add.to(_dev())
    seed = _seed(seed_v)
    drop = (p, salt, seed)
    out, lse = ops.attention_fwd(q, k, v, b, lq, lk, key_mask=key_mask, frames=frames, save_lse=True, drop=drop)
    keep = torch.from_numpy(keep_attn(seed_v, salt, b, 12, lq, lk, p)).to(_dev())
    assert abs(float(keep.float().mean()) - (1 - p)) < 1e-2
    # the keep-bit words the forward hands to the backward are that same mask
    words = lse[2].to(torch.int64) & 0xFFFFFFFF
    kk = torch.arange(lk, device=_dev())
    got_keep = ((words[..., kk >> 5] >> (kk & 31)) & 1).bool()
    assert torch.equal(got_keep, keep)

    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(b, lq, 12, 64).transpose(1, 2)
    kh = kf.view(b, lk, 12, 64).transpose(1, 2)
    vh = vf.view(b, lk, 12, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * 0.125
    if add is not None:
        s = s + add
    pr = torch.softmax(s, -1)
    prd = torch.where(keep, pr / (1 - p), torch.zeros_like(pr))
    ref = (prd @ vh).transpose(1, 2).reshape(b * lq, 768)
    assert (out.float() - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert torch.allclose(lse[0], torch.logsumexp(s, -1), atol=2e-2, rtol=1e-3)     # lse is that of the full softmax

    d_o = torch.randn(b * lq, 768, generator=g).to(torch.bfloat16).to(_dev())
    dq = torch.zeros(b * lq, 768, device=_dev(), dtype=torch.bfloat16)
    dk = torch.zeros(b * lk, 768, device=_dev(), dtype=torch.bfloat16)
    dv = torch.zeros(b * lk, 768, device=_dev(), dtype=torch.bfloat16)
    ops.attention_bwd(q, k, v, out, d_o, lse, dq, dk, dv, b, lq, lk, key_mask=key_mask, frames=frames, drop=drop)
    dq2, dk2, dv2 = torch.zeros_like(dq), torch.zeros_like(dk), torch.zeros_like(dv)
    ops.attention_bwd(q, k, v, out, d_o, lse[0], dq2, dk2, dv2, b, lq, lk, key_mask=key_mask, frames=frames,
                      drop=drop)                                  # two-sweep delta route, mask re-hashed (no bit words)
    dq3, dk3, dv3 = torch.zeros_like(dq), torch.zeros_like(dk), torch.zeros_like(dv)
    ops.attention_bwd(q, k, v, out, d_o, (lse[0], lse[1], None), dq3, dk3, dv3, b, lq, lk, key_mask=key_mask,
                      frames=frames, drop=drop)                   # fp32-O delta route, mask re-hashed (no bit words)
    ref.backward(d_o.float())
    for name, got, got2, got3, want in (("dq", dq, dq2, dq3, qf.grad), ("dk", dk, dk2, dk3, kf.grad),
                                        ("dv", dv, dv2, dv3, vf.grad)):
        tol = 3e-2 * want.abs().max().item() + 1e-6
        assert (got.float() - want).abs().max().item() <= tol, name
        assert (got2.float() - want).abs().max().item() <= tol, name + " (two-sweep)"
        assert (got3.float() - want).abs().max().item() <= tol, name + " (re-hashed mask)"


def test_embedding_dropout(ops):
    from shg_vqa_b200.ops import call, ptr, stream
    g = torch.Generator().manual_seed(8)
    rows, nq, n, p = 2 * 128, 128, 768, 0.1
    word = torch.randn(nq, n, generator=g).to(_dev())
    typ = torch.randn(16, n, generator=g).to(_dev())
    seg = (torch.arange(nq) // 8).repeat(2).to(_dev())
    gamma, beta = torch.ones(n, device=_dev()), torch.zeros(n, device=_dev())
    seed = _seed(31)

    def run(pp):
        y = torch.empty(rows, n, device=_dev(), dtype=torch.bfloat16)
        call("shg_embed_ln_fwd", ptr(word), None, None, ptr(typ), ptr(seg), ptr(gamma), ptr(beta), 1e-12, ptr(y), None,
             None, None, rows, nq, n, pp, 2, ptr(seed) if pp > 0 else None, stream())
        return y

    y0, y1 = run(0.0), run(p)
    keep = torch.from_numpy(keep_flat(31, 2, rows, n, p)).to(_dev())
    ref = torch.where(keep, y0.float() / (1 - p), torch.zeros_like(y0.float()))
    assert (y1.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert torch.equal(y1[~keep], torch.zeros_like(y1[~keep]))


def _trainer(dropout, seed=0):
    from shg_vqa_b200.trainer import HotPathTrainer
    tr = HotPathTrainer(_dev(), dropout=dropout, dropout_seed=seed, t_total=1000)
    tr.init_weights(seed=1)
    return tr


def _batch(b=2):
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import csr_from_padded
    raw = make_inputs(b, seed=3, with_feat=False, visn_tokens=True)
    inp = {k: raw[k].to(_dev()) for k in ("input_ids", "input_mask", "segment_ids", "rel_segment_ids",
                                          "act_segment_ids")}
    inp["visn_tokens"] = raw["visn_tokens"].reshape(-1, 768).to(torch.bfloat16).to(_dev())
    batch = {"target": raw["target"].to(_dev())}
    for fam in ("rel", "act"):
        lab, off = csr_from_padded(raw[fam + "_tokens"], raw[fam + "_lengths"])
        batch[fam + "_labels"], batch[fam + "_off"] = lab.to(_dev()), off.to(_dev())
    return inp, batch


def test_train_step_with_dropout_is_deterministic_and_differs():
    """Whole path, train mode: same (seed, step) → bit-identical losses and gradients (forward and backward draw the
    same masks); a different seed or the next step → a different draw; p → 0 reproduces the no-dropout step."""
    inp, batch = _batch()

    def one(tr, steps=1):
        for _ in range(steps):
            tr.store.grad.zero_()
            tr._advance_seed()
            out, ctx = tr.engine.forward(inp, train=True)
            scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, batch)
            tr.engine.backward(ctx, d_hg, d_rel, d_act)
        torch.cuda.synchronize()
        return scal.clone(), tr.store.grad[tr.store.live_extent].clone(), out

    a_s, a_g, a_out = one(_trainer(True, seed=5))
    b_s, b_g, _ = one(_trainer(True, seed=5))
    assert torch.allclose(a_s[:9], b_s[:9], rtol=1e-5, atol=1e-5)     # loss sums use fp32 atomics
    # split-K wgrads accumulate with fp32 atomics → order noise only
    assert (a_g - b_g).abs().max().item() <= 1e-4 * a_g.abs().max().item()
    c_s, c_g, _ = one(_trainer(True, seed=6))
    assert not torch.equal(a_s[:9], c_s[:9])
    tr = _trainer(True, seed=5)
    d_s, _, _ = one(tr, steps=2)                        # the seed advances every step
    assert not torch.equal(a_s[:9], d_s[:9])
    n_s, n_g, n_out = one(_trainer(False))
    assert not torch.equal(a_s[:9], n_s[:9])
    assert torch.isfinite(a_g).all() and torch.isfinite(a_s).all()
    # dropout perturbs but does not change the scale of the result
    for k in ("rel_logits", "act_logits", "hg_logit"):
        r = float((a_out[k].float() - n_out[k].float()).norm() / n_out[k].float().norm())
        assert 0.0 < r < 1.0, (k, r)
    gn_a, gn_n = float(a_g.norm()), float(n_g.norm())
    assert 0.5 < gn_a / gn_n < 2.0, (gn_a, gn_n)
    # p -> 0: every site keeps everything; the step reproduces the no-dropout one
    tr0 = _trainer(True, seed=5)
    e = tr0.engine
    e.p_hidden = e.p_attn = e.p_dec = e.p_emb = e.p_relq = e.p_actq = 1e-7
    z_s, z_g, _ = one(tr0)
    # (1/(1-p) = 1 + 1 ulp still flips a few bf16 roundings → agreement to bf16 noise, not bitwise)
    assert torch.allclose(z_s[:9], n_s[:9], rtol=2e-3, atol=1e-3)
    assert float((z_g - n_g).norm() / n_g.norm()) <= 5e-2
