"""Pins the oracle (oracle/shg_oracle.py, oracle/lsap.c) against vectors produced by
the UNMODIFIED reference (tools/gen_golden.py) and by the installed scipy."""
import os

import numpy as np
import pytest
import torch

from oracle import shg_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol=2e-4):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b).max() if a.size else 0.0
    scale = max(1.0, np.abs(b).max() if b.size else 1.0)
    assert err <= tol * scale, "max abs err %g (scale %g)" % (err, scale)


def test_inputs_regenerate_identically(golden_hgqa, inputs_b2):
    # the golden file stores checksums of the seeded inputs it was made from
    assert float(inputs_b2["feat"].double().sum()) == float(golden_hgqa["check_feat_sum"])
    assert int(inputs_b2["input_ids"].sum()) == int(golden_hgqa["check_ids_sum"])


@pytest.fixture(scope="module")
def oracle_run(oracle_weights, inputs_b2):
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in
          {so.canonical_key(k): v for k, v in oracle_weights.items()}.items()}
    # expose canonical tensors under every alias name
    full = {k: sd[so.canonical_key(k)] for k in oracle_weights}
    out = so.agqa_forward(full, inputs_b2, task="hgqa")
    loss = so.hgqa_total_loss(out, inputs_b2)
    loss["total"].backward()
    return sd, out, loss


def test_forward_matches_reference(oracle_run, golden_hgqa):
    _, out, _ = oracle_run
    g = golden_hgqa
    _close(out["lang_l"].detach(), g["lang_l"])
    _close(out["visn_r"].detach()[:, ::4], g["visn_r"])
    _close(out["lang_mask"], g["lang_mask"])
    _close(out["logit"].detach(), g["logit"])
    _close(out["rel_logits"].detach(), g["rel_logits"])
    _close(out["act_logits"].detach(), g["act_logits"])
    _close(out["hg_logit"].detach(), g["hg_logit"])


def test_matching_and_loss_match_reference(oracle_run, golden_hgqa):
    _, _, loss = oracle_run
    g = golden_hgqa
    for fam, q_f in (("rel", 8), ("act", 3)):
        for f, (r, c) in enumerate(loss[fam + "_idx"]):
            k = len(r)
            assert np.array_equal(r.numpy(), g[fam + "_rows"][f, :k]) and (g[fam + "_rows"][f, k:] == -1).all()
            assert np.array_equal(c.numpy(), g[fam + "_cols"][f, :k])
    _close(loss["total"].detach(), g["loss_total"], 1e-5)
    _close(loss["qa"].detach(), g["loss_qa"], 1e-5)
    _close(loss["rel_ce"].detach(), g["loss_rel"], 1e-5)
    _close(loss["act_ce"].detach(), g["loss_act"], 1e-5)
    _close(loss["rel_err"], g["rel_err"], 1e-5)
    _close(loss["act_err"], g["act_err"], 1e-5)


def test_gradients_match_reference(oracle_run, golden_hgqa):
    sd, _, _ = oracle_run
    g = golden_hgqa
    names = [str(n) for n in g["grad_names"]]
    for n, ref_norm in zip(names, g["grad_norms"]):
        assert sd[n].grad is not None, n
        mine = sd[n].grad.double().norm().item()
        assert abs(mine - ref_norm) <= 1e-3 * max(ref_norm, 1e-3), (n, mine, ref_norm)
    # tensors the reference leaves without gradient in --taskHGQA training
    for n in g["nograd_names"]:
        t = sd.get(str(n))
        assert t is None or t.grad is None or float(t.grad.abs().max()) == 0.0, n
    tags = {"g_l0_q": "lxrt_encoder.model.bert.encoder.layer.0.attention.self.query.weight",
            "g_r4_ffn": "lxrt_encoder.model.bert.encoder.r_layers.4.intermediate.dense.weight",
            "g_rel_d0_inproj": "rel_decoder.layers.0.self_attn.in_proj_weight",
            "g_act_d4_l2": "action_decoder.layers.4.linear2.weight",
            "g_hgq_q": "hgq_encoder.cross_attn_layer.cross.visual_attention.att.query.weight",
            "g_logit3": "logit_fc.3.weight", "g_cls_embed3_b": "class_embed.3.bias",
            "g_relq_word": "relation_query_embed.word_embeddings.weight",
            "g_word_emb": "lxrt_encoder.model.bert.embeddings.word_embeddings.weight"}
    for tag, key in tags.items():
        mine = sd[key].grad.reshape(-1)[::97][:4096].numpy()
        ref = g[tag]
        assert np.abs(mine - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-6) + 1e-7, tag


def test_vqa_forward_matches_reference(oracle_weights, inputs_b2):
    path = os.path.join(GOLD, "vqa_b2.npz")
    g = np.load(path)
    out = so.agqa_forward(oracle_weights, inputs_b2, task="vqa")
    _close(out["logit"], g["logit"])


def test_lsap_known_answers():
    g = np.load(os.path.join(GOLD, "lsap_kat.npz"))
    for cost, rows, cols, (nr, nc) in zip(g["cost"], g["rows"], g["cols"], g["shapes"]):
        r, c = so.lsap(cost[:nr, :nc])
        k = min(nr, nc)
        assert np.array_equal(r, rows[:k]) and np.array_equal(c, cols[:k])


def test_lsap_fuzz_against_installed_scipy():
    """scipy is the un-vendored third-party dependency the reference calls
    (matcher.py:8); it is present on every box of this image."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(7)
    for case in range(4000):
        nr, nc = int(rng.integers(0, 10)), int(rng.integers(0, 10))
        if case % 3 == 0:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float32)
        elif case % 3 == 1:
            c = -rng.random((nr, nc)).astype(np.float32)
        else:
            base = -rng.random((nr, 3)).astype(np.float32)
            c = base[:, rng.integers(0, 3, nc)]
        r, cc = so.lsap(c)
        r2, c2 = linear_sum_assignment(c)
        assert np.array_equal(r, r2) and np.array_equal(cc, c2), (case, c)


def test_lsap_rejects_invalid_entries_like_scipy():
    with pytest.raises(ValueError):
        so.lsap(np.array([[np.nan, 1.0]], dtype=np.float32))
    with pytest.raises(ValueError):
        so.lsap(np.array([[-np.inf, 1.0]], dtype=np.float32))
    with pytest.raises(ValueError):
        so.lsap(np.array([[np.inf], [np.inf]], dtype=np.float32))   # infeasible
