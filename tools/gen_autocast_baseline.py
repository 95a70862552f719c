"""SURVEY.md §8c yardstick: "per-layer error <= 2x the error of torch.autocast(bf16) on the same module".

Runs the oracle (oracle/shg_oracle.py, the pinned fp32 restatement of the reference) twice on config 1 (B=2, the
inputs and weights of tests/golden/hgqa_b2.npz) on CPU — in fp32 and under torch.autocast("cpu", bfloat16) — with
the reference's assignment held fixed, and writes the relative error of every live tensor's gradient NORM under
autocast (against the golden fp32 norms of the unmodified reference) plus the forward errors to
tests/golden/autocast_bf16_err.json.  tests/test_gpu_model.py bounds the CUDA path's errors by 2x these.

    python tools/gen_autocast_baseline.py        # ~1 min on 8 cores; test infrastructure only
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shg_oracle as so                     # noqa: E402
from shg_vqa_b200.synthetic import make_inputs          # noqa: E402


def run(sd, inp, gold, autocast):
    leaves = {}
    for k, v in sd.items():
        ck = so.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = v.clone().requires_grad_(v.is_floating_point())
    full = {k: leaves[so.canonical_key(k)] for k in sd}
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out = so.agqa_forward(full, inp)
    out = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in out.items()}
    tot = F.binary_cross_entropy_with_logits(out["hg_logit"], inp["target"]) * out["hg_logit"].shape[1]
    for fam in ("rel", "act"):
        rows, cols = gold[fam + "_rows"], gold[fam + "_cols"]
        idx = []
        for f in range(rows.shape[0]):
            k = int((rows[f] >= 0).sum())
            idx.append((torch.from_numpy(rows[f, :k].astype(np.int64)), torch.from_numpy(cols[f, :k].astype(np.int64))))
        tot = tot + so.set_loss(out[fam + "_logits"], inp[fam + "_labels"], idx)[0]
    tot.backward()
    grads = {k: float(v.grad.detach().double().norm()) for k, v in leaves.items() if v.grad is not None}
    return out, grads, float(tot.detach())


def main():
    torch.manual_seed(0)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "hgqa_b2.npz"))
    ks = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    sd = so.make_state_dict(ks, seed=0)
    inp = make_inputs(2, seed=1234)
    out, grads, loss = run(sd, inp, gold, autocast=True)
    ref = dict(zip([str(n) for n in gold["grad_names"]], [float(x) for x in gold["grad_norms"]]))
    res = {"how": "oracle under torch.autocast('cpu', bfloat16), B=2, reference assignment fixed; torch %s" % torch.__version__,
           "loss_total_rel_err": abs(loss - float(gold["loss_total"])) / float(gold["loss_total"]),
           "grad_norm_rel_err": {}, "fwd_rel_l2": {}, "logit_maxabs_over_max": {}}
    for k, v in grads.items():
        if k in ref and ref[k] > 1e-5:
            res["grad_norm_rel_err"][k] = abs(v - ref[k]) / ref[k]
    for name in ("lang_l", "visn_r"):
        mine = out[name].detach()
        r = torch.from_numpy(gold[name])
        if name == "visn_r":
            mine = mine[:, ::4]
        res["fwd_rel_l2"][name] = float((mine.double() - r.double()).norm() / r.double().norm())
    for name in ("logit", "rel_logits", "act_logits", "hg_logit"):
        r = torch.from_numpy(gold[name])
        res["logit_maxabs_over_max"][name] = float((out[name].detach().reshape(r.shape) - r).abs().max() / r.abs().max())
    with open(os.path.join(ROOT, "tests", "golden", "autocast_bf16_err.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    worst = sorted(res["grad_norm_rel_err"].items(), key=lambda kv: -kv[1])[:10]
    print("autocast(bf16) oracle: loss err %.2e; forward %s; logits %s" % (res["loss_total_rel_err"], res["fwd_rel_l2"],
                                                                          res["logit_maxabs_over_max"]))
    for k, v in worst:
        print("  %.4f  %s" % (v, k))


if __name__ == "__main__":
    main()
