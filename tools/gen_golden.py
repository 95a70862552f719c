"""Generate the committed golden vectors under tests/golden/ by running the
UNMODIFIED reference (read-only /root/reference) on CPU.  Build container only.

    python tools/gen_golden.py --task hgqa     # tests/golden/{state_dict_keys.json,hgqa_b2.npz}
    python tools/gen_golden.py --task vqa      # tests/golden/{state_dict_keys_vqa.json,vqa_b2.npz}
    python tools/gen_golden.py --task lsap     # tests/golden/lsap_kat.npz  (scipy known answers)

One task per process: the reference evaluates its argparse singleton at import.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLD = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def gen_model(task):
    import ref_harness as rh
    from oracle import shg_oracle as so
    from shg_vqa_b200.synthetic import make_inputs

    ns = rh.import_reference("--taskHGQA" if task == "hgqa" else "--taskVQA")
    model = rh.build_reference_model(ns)
    sd = model.state_dict()
    keys = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}
    # the alias structure implied by canonical_key must be the reference's own
    groups = {}
    for k, v in sd.items():
        groups.setdefault(v.data_ptr(), []).append(k)
    for ks in groups.values():
        assert len({so.canonical_key(k) for k in ks}) == 1, ks
    assert len({so.canonical_key(k) for k in sd}) == len(groups)
    suffix = "" if task == "hgqa" else "_vqa"
    with open(os.path.join(GOLD, "state_dict_keys%s.json" % suffix), "w") as f:
        json.dump(keys, f, indent=0)

    weights = so.make_state_dict(keys, seed=0)
    model.load_state_dict(weights, strict=True)
    model.eval()   # dropout off; gradients still flow

    inp = make_inputs(2, seed=1234)
    gold = {"check_feat_sum": np.float64(inp["feat"].double().sum().item()),
            "check_ids_sum": np.int64(inp["input_ids"].sum().item())}
    if task == "vqa":
        with torch.no_grad():
            logit, attn = model(inp["feat"], inp["boxes"], inp["input_ids"], inp["input_mask"], inp["segment_ids"],
                                None, None, None, None, None)
        gold["logit"] = _np(logit)
        np.savez_compressed(os.path.join(GOLD, "vqa_b2.npz"), **gold)
        print("vqa logit", logit.abs().mean().item())
        return

    cap = {}
    hook = model.lxrt_encoder.register_forward_hook(lambda m, i, o: cap.update(attn=o[2]))
    logit, rel, act, hg_logit, _ = model(
        inp["feat"], inp["boxes"], inp["input_ids"], inp["input_mask"], inp["segment_ids"],
        inp["rel_segment_ids"], None, inp["act_segment_ids"], None, inp["hg_mask"])
    hook.remove()
    lang_l, lang_mask, visn_r, _ = cap["attn"][-1]     # agqa_model.py:218
    # targets through the reference's own featuriser (entry.py:76-97, agqaHGQA.py:270-297)
    rel_feats = ns.entry.convert_relations_to_features(inp["rel_tokens"], num_rel=8, num_situations=16,
                                                       lengths=inp["rel_lengths"], loss_hg_per_frame=True)
    act_feats = ns.entry.convert_relations_to_features(inp["act_tokens"], num_rel=3, num_situations=16,
                                                       lengths=inp["act_lengths"], loss_hg_per_frame=True)
    tgts = [{"labels": f.targets} for f in rel_feats]
    act_tgts = [{"labels": f.targets} for f in act_feats]
    matcher = ns.matcher.HungarianMatcher(cost_class=1, loss_hg_per_frame=True, clip_len=16)
    A = rh.reference_loss_object(ns)
    w_rel = torch.ones(457); w_rel[0] = 0.1
    w_act = torch.ones(158); w_act[0] = 0.1
    idx = matcher({"pred_logits": rel}, tgts)
    rel_loss = A.loss_labels({"pred_logits": rel}, tgts, idx, empty_weight=w_rel, loss_hg_per_frame=True)
    aidx = matcher({"pred_logits": act}, act_tgts)
    act_loss = A.loss_labels({"pred_logits": act}, act_tgts, aidx, empty_weight=w_act, loss_hg_per_frame=True)
    qa = torch.nn.BCEWithLogitsLoss()(hg_logit, inp["target"]) * hg_logit.size(1)
    total = qa + rel_loss["loss_ce"] + act_loss["loss_ce"]
    total.backward()

    def pack(indices, q_f):
        rows = -np.ones((len(indices), q_f), dtype=np.int64)
        cols = -np.ones((len(indices), q_f), dtype=np.int64)
        for f, (r, c) in enumerate(indices):
            rows[f, :len(r)] = _np(r)
            cols[f, :len(c)] = _np(c)
        return rows, cols

    gold.update(logit=_np(logit), rel_logits=_np(rel), act_logits=_np(act), hg_logit=_np(hg_logit),
                lang_l=_np(lang_l), visn_r=_np(visn_r)[:, ::4].copy(), lang_mask=_np(lang_mask),
                loss_total=_np(total), loss_qa=_np(qa), loss_rel=_np(rel_loss["loss_ce"]),
                loss_act=_np(act_loss["loss_ce"]), rel_err=_np(rel_loss["class_error"]),
                act_err=_np(act_loss["class_error"]))
    gold["rel_rows"], gold["rel_cols"] = pack(idx, 8)
    gold["act_rows"], gold["act_cols"] = pack(aidx, 3)
    # gradients: L2 norm of every parameter that received one (canonical names), and a
    # strided sample of values for a few tensors along the backward path
    names, norms, nograd = [], [], []
    seen = set()
    for k, p in model.named_parameters():          # named_parameters de-duplicates aliases
        ck = so.canonical_key(k)
        if ck in seen:
            continue
        seen.add(ck)
        if p.grad is None:
            nograd.append(ck)
        else:
            names.append(ck)
            norms.append(p.grad.double().norm().item())
    gold["grad_names"] = np.array(names)
    gold["grad_norms"] = np.array(norms)
    gold["nograd_names"] = np.array(nograd)
    params = dict(model.named_parameters())
    for tag, k in [("g_l0_q", "lxrt_encoder.model.bert.encoder.layer.0.attention.self.query.weight"),
                   ("g_r4_ffn", "lxrt_encoder.model.bert.encoder.r_layers.4.intermediate.dense.weight"),
                   ("g_rel_d0_inproj", "rel_decoder.layers.0.self_attn.in_proj_weight"),
                   ("g_act_d4_l2", "action_decoder.layers.4.linear2.weight"),
                   ("g_hgq_q", "hgq_encoder.cross_attn_layer.cross.visual_attention.att.query.weight"),
                   ("g_logit3", "logit_fc.3.weight"),
                   ("g_cls_embed3_b", "class_embed.3.bias"),
                   ("g_relq_word", "relation_query_embed.word_embeddings.weight"),
                   ("g_word_emb", "lxrt_encoder.model.bert.embeddings.word_embeddings.weight")]:
        g = params[k].grad
        gold[tag] = _np(g.reshape(-1)[::97][:4096].clone())
    np.savez_compressed(os.path.join(GOLD, "hgqa_b2.npz"), **gold)
    print("losses", total.item(), qa.item(), rel_loss["loss_ce"].item(), act_loss["loss_ce"].item(),
          rel_loss["class_error"].item(), act_loss["class_error"].item())
    print("n grads", len(names), "no-grad", len(nograd))


def gen_lsap():
    """Known-answer vectors from the installed scipy (the operative third-party oracle)."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(20261018)
    mats, rows, cols, shapes = [], [], [], []
    for case in range(600):
        nr = int(rng.integers(1, 9))
        nc = int(rng.integers(0, 9)) if case % 7 else int(rng.integers(0, 33))
        kind = case % 5
        if kind == 0:
            c = rng.standard_normal((nr, nc)).astype(np.float32)
        elif kind == 1:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float32)          # heavy ties
        elif kind == 2:
            c = np.full((nr, nc), -0.25, dtype=np.float32)                # constant
        elif kind == 3:                                                   # duplicated columns (repeated labels)
            base = -rng.random((nr, max(nc, 1))).astype(np.float32)
            c = base[:, rng.integers(0, max(nc, 1) // 2 + 1, nc)] if nc else base[:, :0]
        else:
            c = -np.exp(rng.standard_normal((nr, nc))).astype(np.float32) / 400.0   # softmax-like magnitudes
        r, cc = linear_sum_assignment(c)
        m = np.zeros((8, 32), dtype=np.float32); m[:nr, :nc] = c
        rr = -np.ones(8, dtype=np.int64); rr[:len(r)] = r
        ccc = -np.ones(8, dtype=np.int64); ccc[:len(cc)] = cc
        mats.append(m); rows.append(rr); cols.append(ccc); shapes.append((nr, nc))
    import scipy
    np.savez_compressed(os.path.join(GOLD, "lsap_kat.npz"), cost=np.stack(mats), rows=np.stack(rows),
                        cols=np.stack(cols), shapes=np.array(shapes, dtype=np.int32),
                        scipy_version=np.array(scipy.__version__))
    print("lsap KATs:", len(mats), "scipy", scipy.__version__)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", required=True, choices=["hgqa", "vqa", "lsap"])
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    if a.task == "lsap":
        gen_lsap()
    else:
        gen_model(a.task)
