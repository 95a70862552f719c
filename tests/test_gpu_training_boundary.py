"""The drop-in boundary TRAINS: `compat.AGQAHotPath` behind `loss.backward()` / torch optimisers.

1. the autograd bridge against HotPathTrainer (same weights, inputs, assignment): losses, gradients, one optimiser step;
2. --taskVQA training backward (LXRT cross layers, pooler, answer head: lxrt/modeling_capsbert.py:658-677,1525-1533)
   against the fp32 oracle's autograd;
3. the reference's UNMODIFIED train loop (tasks/agqaHGQA.py:230-456: its featurisers, its HungarianMatcher call, its
   loss_labels, total_loss.backward(), clip_grad_norm_, BertAdam.step) run once on the reference's own AGQAModel
   (stock PyTorch fp32, CUDA) and once on AGQAHotPath: same weights in, same weights out within the bf16 tolerance.
   Needs the reference's sources: /root/reference in the build container, or a staged copy under baseline/_ref
   (tools/stage_reference.py; git-ignored) on the GPU box — skipped when neither is there.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _reference_style_loss(model_out, raw, matcher, dev):
    """What tasks/agqaHGQA.py:326-378 computes around the model call, restated with torch ops: BCE * n_answers,
    matcher -> target classes (background 0) -> weighted CE (w[0] = 0.1) for relations and actions."""
    logit, rel, act, hg_logit, _ = model_out
    total = F.binary_cross_entropy_with_logits(hg_logit, raw["target"].to(dev)) * hg_logit.shape[1]
    parts = {"qa": float(total.detach())}
    for fam, lg, q_f in (("rel", rel, 8), ("act", act, 3)):
        b = lg.shape[0]
        labels = raw[fam + "_labels"]
        tg = [{"labels": [l.to(dev) for l in labels[i * 16:(i + 1) * 16]]} for i in range(b)]
        idx = matcher({"pred_logits": lg}, tg)
        tgt = torch.zeros(b * 16, q_f, dtype=torch.int64)
        for f, (r, c) in enumerate(idx):
            tgt[f, r] = labels[f][c]
        w = torch.ones(lg.shape[-1], device=dev)
        w[0] = 0.1
        ce = F.cross_entropy(lg.reshape(b * 16, q_f, -1).transpose(1, 2), tgt.to(dev), w)
        parts[fam] = float(ce.detach())
        total = total + ce
    return total, parts


def test_autograd_bridge_matches_trainer(oracle_weights):
    from shg_vqa_b200.compat import AGQAHotPath, HungarianMatcher
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import HotPathTrainer, csr_targets
    dev = torch.device("cuda:0")
    raw = make_inputs(2, seed=31, with_feat=True)
    model = AGQAHotPath(171, num_queries=128, num_classes=456, num_actions=157, dropout=False)
    model.load_state_dict(oracle_weights, strict=True)
    model.train()
    i = {k: v.to(dev) for k, v in raw.items() if torch.is_tensor(v)}
    out = model(i["feat"], i["boxes"], i["input_ids"], i["input_mask"], i["segment_ids"], i["rel_segment_ids"], None,
                i["act_segment_ids"], None, i["hg_mask"])
    assert out[3].requires_grad and out[1].requires_grad
    total, parts = _reference_style_loss(out, raw, HungarianMatcher(1, True, 16), dev)
    params = list(model.parameters())
    assert all(p.grad is None for p in params)
    total.backward()
    named = dict(model.named_parameters())
    # tensors the reference leaves without gradient have .grad None here too (its BertAdam skips them)
    assert named["lxrt_encoder.model.bert.encoder.cross_attn_layer.cross.visual_attention.att.query.weight"].grad is None
    assert named["hgq_encoder.cross_attn_layer.old.visn_inter.dense.weight"].grad is None

    tr = HotPathTrainer(dev, frontend=True, lr=1e-3, t_total=-1)
    tr.load_state_dict(oracle_weights, strict=True)
    inp = {k: i[k] for k in ("feat", "input_ids", "input_mask", "segment_ids", "rel_segment_ids", "act_segment_ids")}
    batch = {"target": i["target"]}
    batch["rel_labels"], batch["rel_off"] = csr_targets(raw["rel_labels"], dev)
    batch["act_labels"], batch["act_off"] = csr_targets(raw["act_labels"], dev)
    o2, ctx = tr.engine.forward(inp, train=True)
    scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(o2, batch)
    tr.engine.backward(ctx, d_hg, d_rel, d_act)
    torch.cuda.synchronize()
    mine = tr.losses_from_scalars(scal)
    for a, b in ((parts["qa"], mine["qa"]), (parts["rel"], mine["rel_ce"]), (parts["act"], mine["act_ce"])):
        assert abs(a - b) <= 1e-4 * abs(b), (a, b)
    worst = 0.0
    for k in ("rel_decoder.layers.2.linear1.weight", "lxrt_encoder.model.bert.encoder.r_layers.0.attention.self.query.weight",
              "lxrt_encoder.model.bert.encoder.layer.3.output.dense.weight", "class_embed.3.weight", "logit_fc.0.bias",
              "lxrt_encoder.model.bert.encoder.visn_fc.conv.4.weight", "action_query_embed.LayerNorm.bias",
              "hgq_encoder.cross_attn_layer.cross.lang_inter.dense.weight",
              "lxrt_encoder.model.bert.embeddings.word_embeddings.weight"):
        g_bridge, g_tr = named[k].grad, tr.store.g(k)
        assert g_bridge is not None and g_bridge.shape == named[k].shape, k
        e = rel_l2(g_bridge, g_tr)
        worst = max(worst, e)
        # same kernels on the same inputs; the loss gradient entering them was rounded to bf16 from torch's fp32
        # autograd on one side and produced in bf16 by the fused loss kernels on the other
        assert e <= 2e-2, (k, e)
    print("autograd bridge vs trainer: losses equal to 1e-4, worst gradient rel-L2 %.2e" % worst)
    # clip_grad_norm_ + an optimiser acting on the Parameter views == the fused clip + BertAdam on the flat buffers
    live = [p for p in params if p.grad is not None]
    torch.nn.utils.clip_grad_norm_(live, 5.0)
    with torch.no_grad():
        for p in live:                               # BertAdam, lxrt/optimization.py:139-172, first step, lr 1e-3
            m, v = 0.1 * p.grad, 0.001 * p.grad * p.grad
            p.add_(m / (v.sqrt() + 1e-6) + 0.01 * p, alpha=-1e-3)
    tr.optimizer_step()
    torch.cuda.synchronize()
    # BertAdam's first step is sign-like (0.1 g / (sqrt(0.001) |g| + 1e-6)): elements whose gradient is noise-level
    # flip between the two runs (their loss gradients differ by a bf16 rounding), so compare the UPDATE's direction
    for k in ("rel_decoder.layers.2.linear1.weight", "class_embed.3.weight",
              "lxrt_encoder.model.bert.encoder.layer.3.output.dense.weight"):
        d_bridge = (named[k].detach() - oracle_weights[k].to(dev)).flatten()
        d_tr = (tr.store.p(k) - oracle_weights[k].to(dev)).flatten()
        cos = float(d_bridge @ d_tr / (d_bridge.norm() * d_tr.norm()))
        assert float(d_tr.abs().max()) > 0 and cos >= 0.98, (k, cos)
    # the next forward sees the updated weights (bf16 shadow refreshed from the master the optimiser wrote)
    model.eval()
    with torch.no_grad():
        out2 = model(i["feat"], i["boxes"], i["input_ids"], i["input_mask"], i["segment_ids"], i["rel_segment_ids"],
                     None, i["act_segment_ids"], None, i["hg_mask"])
    o3 = tr.eval_step(inp)
    assert rel_l2(out2[3], o3["hg_logit"].float().reshape(out2[3].shape)) <= 2e-2
    assert float((out2[3] - out[3].detach()).abs().max()) > 0
    # zero_grad(set_to_none=True) then another backward: fresh gradients, not accumulated ones
    model.train()
    model.zero_grad(set_to_none=True)
    out = model(i["feat"], i["boxes"], i["input_ids"], i["input_mask"], i["segment_ids"], i["rel_segment_ids"], None,
                i["act_segment_ids"], None, i["hg_mask"])
    (out[3].sum() * 0.0 + out[3][0, 0]).backward()
    g1 = named["logit_fc.3.bias"].grad.clone()
    assert abs(float(g1[0]) - 1.0) <= 1e-2 and float(g1[1:].abs().max()) <= 1e-3


def test_vqa_training_backward_vs_oracle(oracle_weights):
    """--taskVQA training (config 4's model in train mode): loss = BCE(logit, target) * 171 through the cross layers,
    pooler and answer head; every live tensor's gradient norm within 5 % of the fp32 oracle's autograd."""
    from oracle import shg_oracle as so
    from shg_vqa_b200.compat import AGQAHotPath
    from shg_vqa_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    raw = make_inputs(2, seed=77, with_feat=True)
    model = AGQAHotPath(171, task="vqa", dropout=False)
    sd = {k: v for k, v in oracle_weights.items() if k in model.key_shapes}
    model.load_state_dict(sd, strict=True)
    model.train()
    i = {k: v.to(dev) for k, v in raw.items() if torch.is_tensor(v)}
    logit, _ = model(i["feat"], i["boxes"], i["input_ids"], i["input_mask"], i["segment_ids"])
    loss = F.binary_cross_entropy_with_logits(logit, i["target"]) * 171
    loss.backward()
    torch.cuda.synchronize()
    leaves = {}
    for k, v in sd.items():
        ck = so.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = v.clone().requires_grad_(v.is_floating_point())
    full = {k: leaves[so.canonical_key(k)] for k in sd}
    ref = so.agqa_forward(full, raw, task="vqa")
    ref_loss = F.binary_cross_entropy_with_logits(ref["logit"], raw["target"]) * 171
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-2 * float(ref_loss)
    named = dict(model.named_parameters())
    worst = []
    for k, t in leaves.items():
        if not t.is_floating_point():
            continue
        if t.grad is None or float(t.grad.norm()) < 1e-6:
            assert named[k].grad is None or float(named[k].grad.norm()) < 5e-2, k
            continue
        assert named[k].grad is not None, k
        a, b = float(named[k].grad.double().norm()), float(t.grad.double().norm())
        worst.append((abs(a - b) / b, k))
    worst.sort(reverse=True)
    print("taskVQA backward: worst gradient-norm deviations", worst[:5])
    assert worst[0][0] <= 5e-2, worst[:5]
    for k in ("lxrt_encoder.model.bert.encoder.cross_attn_layer.cross.visual_attention.att.query.weight",
              "lxrt_encoder.model.bert.pooler_dict.cross.dense2.weight", "logit_fc.0.weight",
              "lxrt_encoder.model.bert.encoder.r_layers.1.output.dense.weight"):
        assert rel_l2(named[k].grad.cpu(), leaves[k].grad) <= 6e-2, k


# ------------------------------------------------------------------------------------------ the reference's own loop
def _reference_ns():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_harness
    if not ref_harness.available():
        pytest.skip("reference sources not present (/root/reference or baseline/_ref)")
    return ref_harness, ref_harness.import_reference("--taskHGQA")


class _Tok:
    """Offline stand-in for BertTokenizer (the real one downloads its vocabulary): deterministic word -> id."""

    def tokenize(self, s):
        return s.lower().split()

    def convert_tokens_to_ids(self, toks):
        return [101 if t == "[CLS]" else 102 if t == "[SEP]" else 1000 + (sum(map(ord, t)) * 7919) % 29000 for t in toks]


class _Eval:
    def evaluateOverall(self, q):
        return 0.0


def _run_reference_loop(ns, model, raw, lr):
    """One epoch of one batch through AGQA.train (tasks/agqaHGQA.py:230-456), unmodified."""
    T = ns.hgqa
    A = object.__new__(T.AGQA)
    dev = torch.device("cuda:0")
    A.model = torch.nn.DataParallel(model, device_ids=[0])
    A.background_idx, A.clip_len = 0, 16
    A.num_situations, A.num_rel, A.num_act, A.num_actions = 16, 8, 3, 157
    A.max_seq_length = 40
    A.tokenizer = _Tok()
    w = torch.ones(457)
    w[0] = 0.1
    A.empty_weight = w
    wa = torch.ones(158)
    wa[0] = 0.1
    A.empty_weight_acts = wa
    A.matcher = ns.matcher.HungarianMatcher(cost_class=1, loss_hg_per_frame=True, clip_len=16)
    A.weight_dict = {"loss_ce": 1, "loss_bbox": 0, "loss_giou": 0}
    A.bce_loss = torch.nn.BCEWithLogitsLoss()
    from src.lxrt.optimization import BertAdam
    A.optim = BertAdam(list(A.model.parameters()), lr=lr, warmup=-1, t_total=-1)
    A.valid_tuple = None
    A.task_q = A.task_vqa = False
    A.task_hgqa = True
    A.task_vhga = False
    A.output = "/tmp"
    A.save = lambda name: None
    b = raw["input_ids"].shape[0]
    words = ["what", "did", "the", "person", "do", "before", "holding", "a", "cup", "after", "sitting", "on", "sofa"]
    sents = [" ".join(words[(j * 3 + k) % len(words)] for k in range(5 + j)) for j in range(b)]
    batch = (list(range(b)), list(range(b)), raw["feat"], raw["boxes"], sents, raw["rel_tokens"], raw["rel_lengths"],
             raw["act_tokens"], raw["hg_mask"], raw["act_lengths"], raw["target"])
    loader = [batch]
    ns.args.epochs, ns.args.tqdm, ns.args.train_sub_set, ns.args.log_freq = 1, False, False, 1
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        A.train((None, loader, _Eval()), None)
    return buf.getvalue()


def test_reference_train_loop_on_hot_path(oracle_weights):
    harness, ns = _reference_ns()
    from shg_vqa_b200.compat import AGQAHotPath
    from shg_vqa_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    raw = make_inputs(2, seed=5, with_feat=True)
    lr = 1e-4
    # (a) the reference's own model, stock PyTorch fp32 on the GPU, dropout probabilities set to 0 for comparability
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref_model = harness.build_reference_model(ns)
    ref_model.load_state_dict(oracle_weights, strict=True)
    for m in ref_model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    ref_model.to(dev)
    log_ref = _run_reference_loop(ns, ref_model, raw, lr)
    sd_ref = {k: v.detach().float().cpu() for k, v in ref_model.state_dict().items()}
    del ref_model
    torch.cuda.empty_cache()
    # (b) the same loop object, the model swapped for the hot path
    hot = AGQAHotPath(171, num_queries=128, num_classes=456, num_actions=157, dropout=False)
    hot.load_state_dict(oracle_weights, strict=True)
    log_hot = _run_reference_loop(ns, hot, raw, lr)
    sd_hot = {k: v.detach().float().cpu() for k, v in hot.state_dict().items()}
    print("reference loop, reference model:\\n" + log_ref.strip())
    print("reference loop, hot path:\\n" + log_hot.strip())

    def losses(log):
        import re
        return [float(x) for x in re.findall(r"(?:Total loss|HGQA loss|Rel loss|Act loss)= *([0-9.]+)", log)]
    lr_, lh = losses(log_ref), losses(log_hot)
    assert len(lr_) >= 4 and len(lr_) == len(lh)
    for a, b in zip(lh, lr_):
        assert abs(a - b) <= 1e-2 * abs(b), (lh, lr_)
    # one BertAdam step moved the same weights the same way: the UPDATE (new - old) agrees in direction and size.
    # BertAdam's first step is m / (sqrt(v) + eps) = sign-like (0.1 g / (sqrt(0.001) |g| + 1e-6)), so elements whose
    # gradient is noise-level flip sign between fp32 and bf16 arithmetic; compare where the reference's step is the
    # saturated +-3.16 lr (|g| well above eps) on tensors with healthy gradients
    checked = 0
    for k in ("logit_fc.3.weight", "class_embed.0.weight", "rel_decoder.layers.4.linear2.weight",
              "lxrt_encoder.model.bert.encoder.r_layers.4.output.dense.weight",
              "hgq_encoder.cross_attn_layer.cross.visual_attention.att.value.weight"):
        d_ref = sd_ref[k] - oracle_weights[k]
        d_hot = sd_hot[k] - oracle_weights[k]
        assert float(d_ref.abs().max()) > 0 and float(d_hot.abs().max()) > 0, k
        cos = float((d_ref.flatten() @ d_hot.flatten()) / (d_ref.norm() * d_hot.norm()))
        print("   update cosine %-80s %.4f" % (k, cos))
        assert cos >= 0.9, (k, cos)
        checked += 1
    # parameters the reference leaves without gradient did not move on either side
    for k in ("lxrt_encoder.model.bert.encoder.cross_attn_layer.old.visn_inter.dense.weight",
              "lxrt_encoder.model.bert.pooler.dense2.weight"):
        assert torch.equal(sd_ref[k], oracle_weights[k]) and torch.equal(sd_hot[k], oracle_weights[k]), k
    assert checked == 5
