"""End-to-end parity of the CUDA hot path (engine.py over libshgvqa.so) against the golden vectors produced
by the UNMODIFIED reference (tests/golden/hgqa_b2.npz, config 1: B=2) and against the oracle.

The path computes in bf16 with fp32 accumulation/statistics; the reference is fp32.  Tolerances (stated per
assert) follow SURVEY.md §8c: activations rel-L2 <= 2e-2 after 5-12 stacked post-LN layers, logits max-abs <=
3e-2 * max|logit|, losses rel <= 1e-2, gradient norms rel <= 5e-2 (bf16 backward through ~20 layers).
The Conv3d front-end is outside the path (stock PyTorch): the visual tokens fed to the engine are produced by
the oracle's fp32 front-end from the same seeded features the golden run used."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.fixture(scope="module")
def autocast_err():
    """Errors of the oracle run under torch.autocast(bf16) against the same golden vectors (SURVEY.md §8c's
    yardstick: "<= 2x the error of torch.autocast(bf16) on the same module"); tools/gen_autocast_baseline.py."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "autocast_bf16_err.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def run(oracle_weights, inputs_b2, golden_hgqa):
    from oracle import shg_oracle as so
    from shg_vqa_b200.trainer import HotPathTrainer, csr_targets
    dev = torch.device("cuda:0")
    tr = HotPathTrainer(dev)
    tr.load_state_dict(oracle_weights, strict=True)
    tokens = so.visual_front_end(oracle_weights, "lxrt_encoder.model.bert.encoder.visn_fc", inputs_b2["feat"])
    inp = {k: inputs_b2[k].to(dev) for k in ("input_ids", "input_mask", "segment_ids", "rel_segment_ids",
                                             "act_segment_ids")}
    inp["visn_tokens"] = tokens.reshape(-1, 768).to(torch.bfloat16).to(dev)
    batch = {"target": inputs_b2["target"].to(dev)}
    batch["rel_labels"], batch["rel_off"] = csr_targets(inputs_b2["rel_labels"], dev)
    batch["act_labels"], batch["act_off"] = csr_targets(inputs_b2["act_labels"], dev)
    out, ctx = tr.engine.forward(inp, train=True)
    # 1) the path's own matching (reported against the reference's below)
    _, _, _, _, matches = tr.loss_and_grads(out, batch, want_grad=False)
    torch.cuda.synchronize()
    own = [{k: v.clone() for k, v in m.items()} for m in matches]
    # 2) loss + backward with the reference's assignment held fixed, so that gradients are comparable:
    #    at this near-flat operating point bf16 noise in the logits legitimately flips near-tied assignments
    for fam, q_f, labels in (("rel", 8, inputs_b2["rel_labels"]), ("act", 3, inputs_b2["act_labels"])):
        rows, cols = golden_hgqa[fam + "_rows"], golden_hgqa[fam + "_cols"]
        tgt = torch.zeros(rows.shape[0], q_f, dtype=torch.int32)
        for f in range(rows.shape[0]):
            k = int((rows[f] >= 0).sum())
            tgt[f, torch.from_numpy(rows[f, :k])] = labels[f][torch.from_numpy(cols[f, :k])].to(torch.int32)
        batch[fam + "_target_classes"] = tgt.reshape(-1).to(dev)
    scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, batch)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    return tr, out, scal.clone(), own


def test_forward_activations_and_logits(run, golden_hgqa, autocast_err):
    tr, out, _, _ = run
    g = golden_hgqa
    lang_l = out["lang_l"].float().cpu().view(2, 40, 768)
    visn_r = out["visn_r"].float().cpu().view(2, 393, 768)
    e_l, e_v = rel_l2(lang_l, torch.from_numpy(g["lang_l"])), rel_l2(visn_r[:, ::4], torch.from_numpy(g["visn_r"]))
    print("activation rel-L2 vs the reference: lang_l %.4f (autocast bf16: %.4f), visn_r %.4f (autocast: %.4f)"
          % (e_l, autocast_err["fwd_rel_l2"]["lang_l"], e_v, autocast_err["fwd_rel_l2"]["visn_r"]))
    assert e_l <= 2e-2 and e_v <= 2e-2
    # no worse than 2x what torch.autocast(bf16) does to the same stack
    assert e_l <= 2 * autocast_err["fwd_rel_l2"]["lang_l"] and e_v <= 2 * autocast_err["fwd_rel_l2"]["visn_r"]
    for name in ("logit", "rel_logits", "act_logits", "hg_logit"):
        mine = out[name].float().cpu().reshape(g[name].shape)
        ref = torch.from_numpy(g[name])
        err = float((mine - ref).abs().max())
        assert err <= 3e-2 * float(ref.abs().max()), (name, err, float(ref.abs().max()))
        assert rel_l2(mine, ref) <= 3e-2, (name, rel_l2(mine, ref))


def test_matching_and_losses(run, golden_hgqa):
    tr, out, scal, matches = run
    g = golden_hgqa
    losses = tr.losses_from_scalars(scal)
    for mine, ref in ((losses["qa"], g["loss_qa"]), (losses["rel_ce"], g["loss_rel"]),
                      (losses["act_ce"], g["loss_act"]), (losses["total"], g["loss_total"])):
        assert abs(mine - float(ref)) <= 1e-2 * abs(float(ref)), (mine, float(ref))
    # end-to-end matcher agreement (our bf16 logits → our matcher vs reference fp32 logits → scipy) is REPORTED,
    # not required: the logits are near-flat here and agree only to bf16 precision, so near-tied assignments flip.
    # Exactness on identical logits is test_matcher_exact_on_reference_logits; on identical cost matrices,
    # tests/test_gpu_kernels.py::test_lsap_*.
    for fam, m in zip(("rel", "act"), matches):
        rows, cols, n = m["rows"].cpu().numpy(), m["cols"].cpu().numpy(), m["n_match"].cpu().numpy()
        same = 0
        for f in range(rows.shape[0]):
            k = n[f]
            assert (g[fam + "_rows"][f, k:] == -1).all()          # matched counts are exact
            same += np.array_equal(rows[f, :k], g[fam + "_rows"][f, :k]) and \
                np.array_equal(cols[f, :k], g[fam + "_cols"][f, :k])
        print("%s matcher agreement with the reference: %d / %d frames" % (fam, same, rows.shape[0]))


def test_matcher_exact_on_reference_logits(golden_hgqa, inputs_b2):
    """Same fp32 logits as the reference → identical assignments from the fused device matcher."""
    from shg_vqa_b200 import ops
    from shg_vqa_b200.trainer import csr_targets
    dev = torch.device("cuda:0")
    g = golden_hgqa
    for fam, q_f in (("rel", 8), ("act", 3)):
        lg = torch.from_numpy(g[fam + "_logits"]).to(dev)
        lab, off = csr_targets(inputs_b2[fam + "_labels"], dev)
        m = ops.match_from_logits(lg.view(-1, lg.shape[-1]), lab, off, q_f)
        rows, cols, n = m["rows"].cpu().numpy(), m["cols"].cpu().numpy(), m["n_match"].cpu().numpy()
        for f in range(rows.shape[0]):
            k = n[f]
            assert np.array_equal(rows[f, :k], g[fam + "_rows"][f, :k])
            assert np.array_equal(cols[f, :k], g[fam + "_cols"][f, :k])


def test_gradients(run, golden_hgqa, autocast_err):
    tr, _, _, _ = run
    ac = autocast_err["grad_norm_rel_err"]
    g = golden_hgqa
    st = tr.store
    names = [str(n) for n in g["grad_names"]]
    worst = []
    for n, ref in zip(names, g["grad_norms"]):
        if ".visn_fc." in n:
            continue                                   # front-end: outside the path
        mine = float(st.g(n).double().norm())
        if n.endswith(".key.bias"):
            # softmax is invariant to a key bias: the exact gradient is 0 (the reference holds ~1e-7 of fp32
            # round-off); ours is the bf16 round-off of the same sum — bounded absolutely, not relatively
            assert ref < 1e-5 and mine < 5e-2, (n, mine, ref)
            continue
        if n.endswith("layers.0.self_attn.in_proj_weight"):
            # layer 0 attends over a constant value (tgt = 0, tasks/agqa_model.py:235): the exact q/k gradient is 0
            assert ref < 1e-5 and mine < 1e-5, (n, mine, ref)
            continue
        if n.startswith(("relation_query_embed.", "action_query_embed.")):
            # These 8 tensors are fed ONLY by the gradient w.r.t. the per-query position input of the decoders,
            # i.e. by differences BETWEEN queries.  At this (random-init-like) operating point the decoder states
            # of different queries differ by 0.5 % (measured: |V - mean V| / |V| = 0.0055 in the self-attention
            # values), which is the resolution of bf16 (2^-8), so that signal is quantisation-dominated for ANY bf16
            # activation path: the oracle itself under torch.autocast(bf16) is off by 62 % / 31 % / 21 % / 13 % /
            # 12 % / 6 % on them and by <= 1.3 % on every other tensor (tests/golden/autocast_bf16_err.json).
            # SURVEY §8c's rule is the bound: no worse than 2x the autocast error (floor: the 5 % of the rest).
            # Measured (profiles/parity_report_r02.log): 0.79 / 0.61 / 0.27 / 0.17 / 0.13 / 0.07 — 1.0-2.0x autocast's.
            err = abs(mine - ref) / ref
            bound = max(2.0 * ac[n] + 0.02, 5e-2)        # + 0.02: run-to-run slack of a quantisation-dominated quantity
            print("   query-embed gradient %-55s err %.3f  autocast(bf16) %.3f  bound %.3f" % (n, err, ac[n], bound))
            assert err <= bound, (n, mine, ref, err, ac[n])
            continue
        rel = abs(mine - ref) / max(ref, 1e-6)
        worst.append((rel, n, mine, float(ref)))
    worst.sort(reverse=True)
    print("worst gradient-norm deviations:")
    for w in worst[:12]:
        print("   %.4f  %-90s %.5g vs %.5g" % w)
    # bf16 backward through up to ~20 stacked layers: every tensor's gradient norm within 5 %
    assert worst[0][0] <= 5e-2, worst[:5]
    tags = {"g_l0_q": "lxrt_encoder.model.bert.encoder.layer.0.attention.self.query.weight",
            "g_r4_ffn": "lxrt_encoder.model.bert.encoder.r_layers.4.intermediate.dense.weight",
            "g_act_d4_l2": "action_decoder.layers.4.linear2.weight",
            "g_hgq_q": "hgq_encoder.cross_attn_layer.cross.visual_attention.att.query.weight",
            "g_logit3": "logit_fc.3.weight", "g_cls_embed3_b": "class_embed.3.bias",
            "g_word_emb": "lxrt_encoder.model.bert.embeddings.word_embeddings.weight"}
    for tag, key in tags.items():
        mine = st.g(key).reshape(-1)[::97][:4096].float().cpu()
        ref = torch.from_numpy(g[tag])
        assert rel_l2(mine, ref) <= 6e-2, (tag, rel_l2(mine, ref))
    # tensors the reference leaves without gradient stay exactly zero
    for n in g["nograd_names"]:
        assert float(st.g(str(n)).abs().max()) == 0.0, n


def test_state_dict_round_trip_on_device(run, oracle_weights):
    tr, _, _, _ = run
    sd = tr.state_dict()
    assert list(sd) == list(oracle_weights)
    for k in ("logit_fc.3.weight", "rel_decoder.layers.3.multihead_attn.in_proj_weight",
              "hgq_encoder.x_layers.1.visn_output.dense.weight"):
        assert torch.equal(sd[k].cpu(), oracle_weights[k])


def test_optimizer_step_matches_bert_adam(run):
    """clip_grad_norm_(5.) + BertAdam (lxrt/optimization.py:101-180) on the flat buffers vs the same arithmetic
    in plain torch on a copy."""
    tr, _, _, _ = run
    st = tr.store
    live = st.live_extent
    p0 = st.master[live].clone()
    g0 = st.grad[live].clone()
    tr.lr, tr.t_total = 1e-3, -1
    tr.optimizer_step()
    torch.cuda.synchronize()
    norm = g0.double().norm().float()
    gc = g0 * torch.clamp(5.0 / (norm + 1e-6), max=1.0)
    m = 0.1 * gc
    v = 0.001 * gc * gc
    ref = p0 - 1e-3 * (m / (v.sqrt() + 1e-6) + 0.01 * p0)
    assert torch.allclose(st.master[live], ref, rtol=1e-5, atol=1e-7)
    assert float(st.grad[live].abs().max()) == 0.0                      # zeroed for the next step
    # bf16 shadow refreshed from the new master
    n_sh = min(st.n_matrix, st.live_end) - st.live_begin
    assert torch.equal(st.shadow[st.live_begin:st.live_begin + n_sh],
                       st.master[st.live_begin:st.live_begin + n_sh].to(torch.bfloat16))


def test_compat_matcher_returns_reference_structure(golden_hgqa, inputs_b2):
    """HungarianMatcher mirror: same call contract as lxrt/matcher.py (dict in, list of (int64, int64) CPU
    tensors out), identical assignments on the reference's own logits."""
    from shg_vqa_b200.compat import HungarianMatcher
    dev = torch.device("cuda:0")
    g = golden_hgqa
    m = HungarianMatcher(cost_class=1, loss_hg_per_frame=True, clip_len=16)
    for fam, q_f in (("rel", 8), ("act", 3)):
        labels = inputs_b2[fam + "_labels"]
        targets = [{"labels": [l.to(dev) for l in labels[b * 16:(b + 1) * 16]]} for b in range(2)]
        res = m({"pred_logits": torch.from_numpy(g[fam + "_logits"]).to(dev)}, targets)
        assert len(res) == 32
        for f, (r, c) in enumerate(res):
            assert r.dtype == torch.int64 and c.dtype == torch.int64 and not r.is_cuda
            k = len(r)
            assert np.array_equal(r.numpy(), g[fam + "_rows"][f, :k]) and (g[fam + "_rows"][f, k:] == -1).all()
            assert np.array_equal(c.numpy(), g[fam + "_cols"][f, :k])
    with pytest.raises(AssertionError):
        HungarianMatcher(cost_class=0)


def test_compat_model_forward_from_backbone_features(oracle_weights, inputs_b2, golden_hgqa):
    """AGQAModel-shaped forward (stock-PyTorch Conv3d front-end + engine) from the same seeded backbone features
    the reference consumed: the 5-tuple of tasks/agqa_model.py:269 within the bf16 tolerance."""
    from shg_vqa_b200.compat import AGQAHotPath
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    model = AGQAHotPath(171, num_queries=128, num_classes=456, num_actions=157).eval()
    model.load_state_dict(oracle_weights, strict=True)
    i = {k: v.to(dev) for k, v in inputs_b2.items() if torch.is_tensor(v)}
    logit, rel, act, hg_logit, attn = model(i["feat"], i["boxes"], i["input_ids"], i["input_mask"], i["segment_ids"],
                                            i["rel_segment_ids"], None, i["act_segment_ids"], None, i["hg_mask"])
    g = golden_hgqa
    for name, mine in (("logit", logit), ("rel_logits", rel), ("act_logits", act), ("hg_logit", hg_logit)):
        ref = torch.from_numpy(g[name])
        assert tuple(mine.shape) == tuple(ref.shape)
        assert float((mine.float().cpu() - ref).abs().max()) <= 3e-2 * float(ref.abs().max()), name
    assert len(attn) == 5
    assert list(model.state_dict()) == list(oracle_weights)


def test_vqa_eval_forward_config4(oracle_weights, inputs_b2):
    """Config 4 (--taskVQA eval: LXRT stack + pooler + answer head only, no decoders / matcher) against the golden
    logits of the unmodified reference built with --taskVQA (tests/golden/vqa_b2.npz)."""
    import os
    from oracle import shg_oracle as so
    from shg_vqa_b200.compat import AGQAHotPath
    dev = torch.device("cuda:0")
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vqa_b2.npz"))
    model = AGQAHotPath(171, task="vqa").eval()
    sd = {k: v for k, v in oracle_weights.items() if k in model.key_shapes}
    assert len(sd) == 324
    model.load_state_dict(sd, strict=True)
    i = {k: v.to(dev) for k, v in inputs_b2.items() if torch.is_tensor(v)}
    logit, attn = model(i["feat"], i["boxes"], i["input_ids"], i["input_mask"], i["segment_ids"])
    ref = torch.from_numpy(gold["logit"])
    assert float((logit.float().cpu() - ref).abs().max()) <= 3e-2 * float(ref.abs().max())


def test_star_shaped_decoder_matcher_stress_config5():
    """Config 5: STAR class counts (564 relation / 112 action classes, 4 answers), bs64, every frame full
    (1024 8x8 + 1024 3x3 assignment problems): forward + device matching vs the oracle on the same weights;
    assignments compared on the path's own logits fed to the oracle's CPU matcher (bit-exact)."""
    from oracle import shg_oracle as so
    from shg_vqa_b200.params import build_key_shapes
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import HotPathTrainer, csr_targets
    dev = torch.device("cuda:0")
    tr = HotPathTrainer(dev, n_answers=4, n_rel_cls=564, n_act_cls=112)
    ks = {k: [list(s), d] for k, (s, d) in tr.key_shapes.items() if ".visn_fc.conv" not in k}
    w = so.make_state_dict(ks, seed=5)
    tr.load_state_dict(w, strict=False)
    B = 64
    raw = make_inputs(B, seed=77, n_rel_cls=564, n_act_cls=112, n_ans=4, full_frames=True, with_feat=False,
                      visn_tokens=True)
    inp = {k: raw[k].to(dev) for k in ("input_ids", "input_mask", "segment_ids", "rel_segment_ids", "act_segment_ids")}
    inp["visn_tokens"] = raw["visn_tokens"].reshape(-1, 768).to(torch.bfloat16).to(dev)
    batch = {"target": raw["target"].to(dev)}
    batch["rel_labels"], batch["rel_off"] = csr_targets(raw["rel_labels"], dev)
    batch["act_labels"], batch["act_off"] = csr_targets(raw["act_labels"], dev)
    out, ctx = tr.engine.forward(inp, train=True)
    scal, d_hg, d_rel, d_act, matches = tr.loss_and_grads(out, batch)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    assert out["rel_logits"].shape == (B * 128, 564) and out["act_logits"].shape == (B * 48, 112)
    assert out["hg_logit"].shape == (B, 4)
    for fam, q_f, m in (("rel", 8, matches[0]), ("act", 3, matches[1])):
        lg = out[fam + "_logits"].float().cpu().reshape(B, 16 * q_f, -1)
        ref = so.hungarian_match(lg, raw[fam + "_labels"])
        rows, cols, n = m["rows"].cpu().numpy(), m["cols"].cpu().numpy(), m["n_match"].cpu().numpy()
        assert (n == q_f).all() and (m["status"].cpu().numpy() == 0).all()
        same = sum(np.array_equal(rows[f], r.numpy()) and np.array_equal(cols[f], c.numpy())
                   for f, (r, c) in enumerate(ref))
        assert same >= 0.995 * len(ref), "%s: %d / %d" % (fam, same, len(ref))
    losses = tr.losses_from_scalars(scal)
    assert np.isfinite(losses["total"]) and float(tr.store.grad[tr.store.live_extent].abs().max()) > 0
    # oracle forward on 2 of the 64 clips: logits agree within the bf16 tolerance
    sub = {k: (v[:2] if torch.is_tensor(v) else v) for k, v in raw.items()}
    o = so.agqa_forward(w, sub)
    mine = out["rel_logits"].float().cpu().reshape(B, 128, 564)[:2]
    assert float((mine - o["rel_logits"]).abs().max()) <= 3e-2 * float(o["rel_logits"].abs().max())


def test_answer_argmax_agreement_bs32(oracle_weights):
    """north_star: "answer-argmax and loss agreement reported".  B = 32 synthetic clips through the CUDA path and
    through the fp32 oracle (CPU): top-1 agreement of `hg_logit` and `logit`, overall and on the samples whose fp32
    top-2 margin exceeds twice the logit tolerance (3e-2 * max|logit|) — there it must be 100 %; on the
    rest (random-init logits are near-flat) a flip is within tolerance and the rate is only reported."""
    from oracle import shg_oracle as so
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import HotPathTrainer
    dev = torch.device("cuda:0")
    tr = HotPathTrainer(dev, frontend=True)
    tr.load_state_dict(oracle_weights, strict=True)
    raw = make_inputs(32, seed=4321, with_feat=True)
    inp = {k: raw[k].to(dev) for k in ("feat", "input_ids", "input_mask", "segment_ids", "rel_segment_ids",
                                       "act_segment_ids")}
    out = tr.eval_step(inp)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = so.agqa_forward(oracle_weights, raw)
    for name in ("hg_logit", "logit"):
        r = ref[name]
        mine = out[name].float().cpu().reshape(r.shape)
        tol = 3e-2 * float(r.abs().max())
        assert float((mine - r).abs().max()) <= tol, name
        top2 = r.topk(2, dim=1).values
        margin = top2[:, 0] - top2[:, 1]
        agree = mine.argmax(1) == r.argmax(1)
        clear = margin > 2 * tol
        print("%s: answer-argmax agreement %d / %d overall; %d / %d on samples with fp32 top-2 margin > 2 x tol (%.4f)"
              % (name, int(agree.sum()), len(agree), int(agree[clear].sum()), int(clear.sum()), 2 * tol))
        assert bool(agree[clear].all())
    # loss agreement at B = 32 (own matching on each side)
    from shg_vqa_b200.trainer import csr_targets
    batch = {"target": raw["target"].to(dev)}
    batch["rel_labels"], batch["rel_off"] = csr_targets(raw["rel_labels"], dev)
    batch["act_labels"], batch["act_off"] = csr_targets(raw["act_labels"], dev)
    scal, _, _, _, _ = tr.loss_and_grads(out, batch, want_grad=False)
    mine = tr.losses_from_scalars(scal)
    want = so.hgqa_total_loss(ref, raw)
    for k in ("qa", "rel_ce", "act_ce", "total"):
        w = float(want[k])
        print("loss %-6s cuda %.5f  oracle %.5f  rel err %.2e" % (k, mine[k], w, abs(mine[k] - w) / abs(w)))
        assert abs(mine[k] - w) <= 1e-2 * abs(w), (k, mine[k], w)


def test_matcher_end_to_end_on_peaked_logits():
    """A trained-like operating point: every frame's targets have a clearly preferred query (logit margin >> bf16
    resolution).  bf16 logits -> device matcher must then agree with fp32 logits -> the oracle's scipy-equivalent
    solver on >= 99 % of the frames (at the random-init point of the golden run, near-tied costs flip legitimately
    and agreement is only reported, test_matching_and_losses)."""
    from oracle import shg_oracle as so
    from shg_vqa_b200 import ops
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import csr_targets
    dev = torch.device("cuda:0")
    B = 16
    raw = make_inputs(B, seed=99, with_feat=False, visn_tokens=False)
    g = torch.Generator().manual_seed(5)
    for fam, q_f, n_cls in (("rel", 8, 457), ("act", 3, 158)):
        labels = raw[fam + "_labels"]
        lg = torch.randn(B * 16, q_f, n_cls, generator=g)
        for f, lab in enumerate(labels):
            perm = torch.randperm(q_f, generator=g)[:len(lab)]
            for j, c in enumerate(lab.tolist()):
                lg[f, perm[j], c] += 6.0 + 2.0 * float(torch.rand((), generator=g))
        ref = so.hungarian_match(lg.view(B, 16 * q_f, n_cls), labels)
        noisy = (lg * (1.0 + 2.0 ** -8 * torch.randn(lg.shape, generator=g))).to(torch.bfloat16).to(dev)
        lab, off = csr_targets(labels, dev)
        m = ops.match_from_logits(noisy.view(-1, n_cls), lab, off, q_f)
        rows, cols, n = m["rows"].cpu().numpy(), m["cols"].cpu().numpy(), m["n_match"].cpu().numpy()
        same = sum(np.array_equal(rows[f, :n[f]], r.numpy()) and np.array_equal(cols[f, :n[f]], c.numpy())
                   for f, (r, c) in enumerate(ref))
        print("%s: %d / %d frames identical on peaked logits" % (fam, same, len(ref)))
        assert same >= 0.99 * len(ref), (fam, same, len(ref))


def test_matcher_flags_bad_labels_and_nan(golden_hgqa, inputs_b2):
    """A label >= n_cls (IndexError in the reference, matcher.py:73) and NaN logits (scipy ValueError) are never
    dereferenced / never silently trained on: status -4 / -2, and the loss read raises."""
    from shg_vqa_b200 import ops
    from shg_vqa_b200.compat import HungarianMatcher
    from shg_vqa_b200.trainer import HotPathTrainer, csr_targets
    dev = torch.device("cuda:0")
    lg = torch.from_numpy(golden_hgqa["rel_logits"]).to(dev)
    labels = [l.clone() for l in inputs_b2["rel_labels"]]
    f_bad = next(i for i, l in enumerate(labels) if len(l) > 0)
    labels[f_bad][0] = 457                                      # one past the last class
    lab, off = csr_targets(labels, dev)
    m = ops.match_from_logits(lg.view(-1, 457), lab, off, 8)
    st = m["status"].cpu()
    assert int(st[f_bad]) == -4 and int((st != 0).sum()) == 1
    scal = torch.zeros(16, device=dev)
    ops.set_loss_fwd(lg.view(-1, 457), m, 0.1, out=scal[1:5], err_out=scal[10:11])
    with pytest.raises(ValueError):
        HotPathTrainer.losses_from_scalars(scal)
    targets = [{"labels": [l.to(dev) for l in labels[b * 16:(b + 1) * 16]]} for b in range(2)]
    with pytest.raises(IndexError):
        HungarianMatcher(1, True, 16)({"pred_logits": lg}, targets)
    # NaN logits in one frame
    lg2 = lg.clone()
    lg2.view(-1, 457)[f_bad * 8 + 1, 5] = float("nan")
    lab, off = csr_targets(inputs_b2["rel_labels"], dev)
    m = ops.match_from_logits(lg2.view(-1, 457), lab, off, 8)
    assert int(m["status"].cpu()[f_bad]) == -2
    targets = [{"labels": [l.to(dev) for l in inputs_b2["rel_labels"][b * 16:(b + 1) * 16]]} for b in range(2)]
    with pytest.raises(ValueError):
        HungarianMatcher(1, True, 16)({"pred_logits": lg2}, targets)
