"""Ten BertAdam steps against the oracle ("no multi-step check against the oracle").

Config 1 (B = 2), dropout off, constant lr: the CUDA path (forward, device matching, loss, backward, fused
clip_grad_norm_(5.) + BertAdam, lxrt/optimization.py:139-172) and the fp32 oracle on the CPU (torch autograd +
the same update written with torch ops) start from the same weights and see the same batch ten times.  Per step the
total loss must agree to 1e-2 (SURVEY §8c's loss tolerance); both trajectories must go down."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ten_step_loss_trajectory_vs_oracle(oracle_weights, inputs_b2):
    from oracle import shg_oracle as so
    from shg_vqa_b200.trainer import HotPathTrainer, csr_targets
    dev = torch.device("cuda:0")
    lr, steps = 5e-6, 10
    # ---------------------------------------------------------------- CUDA path
    tr = HotPathTrainer(dev, frontend=True, lr=lr, t_total=-1, dropout=False)
    tr.load_state_dict(oracle_weights, strict=True)
    raw = inputs_b2
    inp = {k: raw[k].to(dev) for k in ("feat", "input_ids", "input_mask", "segment_ids", "rel_segment_ids",
                                       "act_segment_ids")}
    batch = {"target": raw["target"].to(dev)}
    batch["rel_labels"], batch["rel_off"] = csr_targets(raw["rel_labels"], dev)
    batch["act_labels"], batch["act_off"] = csr_targets(raw["act_labels"], dev)
    mine = []
    for _ in range(steps):
        _, scal = tr.train_step(inp, batch)
        mine.append(tr.losses_from_scalars(scal))
    # ---------------------------------------------------------------- oracle, CPU fp32
    torch.set_num_threads(os.cpu_count() or 1)
    leaves = {}
    for k, v in oracle_weights.items():
        ck = so.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = v.clone().requires_grad_(v.is_floating_point())
    full = {k: leaves[so.canonical_key(k)] for k in oracle_weights}
    state = {}
    want = []
    for _ in range(steps):
        for t in leaves.values():
            t.grad = None
        out = so.agqa_forward(full, raw)
        L = so.hgqa_total_loss(out, raw)
        L["total"].backward()
        want.append({k: float(L[k]) for k in ("total", "qa", "rel_ce", "act_ce")})
        live = [t for t in leaves.values() if t.grad is not None]
        torch.nn.utils.clip_grad_norm_(live, 5.0)
        with torch.no_grad():
            for t in live:
                m, v = state.setdefault(id(t), (torch.zeros_like(t), torch.zeros_like(t)))
                m.mul_(0.9).add_(t.grad, alpha=0.1)
                v.mul_(0.999).addcmul_(t.grad, t.grad, value=0.001)
                t.add_(m / (v.sqrt() + 1e-6) + 0.01 * t, alpha=-lr)
    for i, (a, b) in enumerate(zip(mine, want)):
        print("step %2d  total: cuda %.4f  oracle %.4f | qa %.4f / %.4f  rel %.4f / %.4f  act %.4f / %.4f"
              % (i, a["total"], b["total"], a["qa"], b["qa"], a["rel_ce"], b["rel_ce"], a["act_ce"], b["act_ce"]))
    for i, (a, b) in enumerate(zip(mine, want)):
        assert abs(a["total"] - b["total"]) <= 1e-2 * abs(b["total"]), (i, a, b)
    assert mine[-1]["total"] < mine[0]["total"] and want[-1]["total"] < want[0]["total"]
    # the two trajectories moved by a comparable amount (not just "both flat")
    d_mine, d_want = mine[0]["total"] - mine[-1]["total"], want[0]["total"] - want[-1]["total"]
    assert abs(d_mine - d_want) <= 0.25 * abs(d_want) + 1e-2 * want[0]["total"], (d_mine, d_want)
