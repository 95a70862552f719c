import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import shg_oracle as so
from shg_vqa_b200.synthetic import make_inputs
from shg_vqa_b200.trainer import HotPathTrainer, csr_targets
ks = json.load(open('tests/golden/state_dict_keys.json'))
W = so.make_state_dict(ks, 0)
g = dict(np.load('tests/golden/hgqa_b2.npz'))
inp = make_inputs(2, seed=1234)
dev = torch.device('cuda:0')
tr = HotPathTrainer(dev); tr.load_state_dict(W)
tokens = so.visual_front_end(W, "lxrt_encoder.model.bert.encoder.visn_fc", inp["feat"])
di = {k: inp[k].to(dev) for k in ("input_ids", "input_mask", "segment_ids", "rel_segment_ids", "act_segment_ids")}
di["visn_tokens"] = tokens.reshape(-1, 768).to(torch.bfloat16).to(dev)
batch = {"target": inp["target"].to(dev)}
batch["rel_labels"], batch["rel_off"] = csr_targets(inp["rel_labels"], dev)
batch["act_labels"], batch["act_off"] = csr_targets(inp["act_labels"], dev)
for fam, q_f in (("rel", 8), ("act", 3)):
    rows, cols = g[fam + "_rows"], g[fam + "_cols"]
    tgt = torch.zeros(rows.shape[0], q_f, dtype=torch.int32)
    for f in range(rows.shape[0]):
        k = int((rows[f] >= 0).sum())
        tgt[f, torch.from_numpy(rows[f, :k])] = inp[fam + "_labels"][f][torch.from_numpy(cols[f, :k])].to(torch.int32)
    batch[fam + "_target_classes"] = tgt.reshape(-1).to(dev)
tr.engine._trace = []
out, ctx = tr.engine.forward(di, train=True)
scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, batch)
tr.engine.backward(ctx, d_hg, d_rel, d_act)
torch.cuda.synchronize()
dbg = {k: v.float().cpu() for k, v in tr.engine._dbg.items()}
# oracle truth: hook the hg-embedding outputs
orig = so.hg_embeddings
leaves = {}
def hooked(sd, p, seg):
    y = orig(sd, p, seg).detach().requires_grad_(True)
    leaves[p] = y
    return y
so.hg_embeddings = hooked
cap = []
orig_dl = so.decoder_layer
import torch.nn.functional as F
def dl(sd, p, tgt, memory, tgt_mask, query_pos):
    qk = tgt + query_pos
    qk.retain_grad()
    t = so._ln(sd, p + ".norm1", tgt + so.mha(sd, p + ".self_attn", qk, qk, tgt, tgt_mask), so.LN_EPS_DETR)
    cq = t + query_pos
    cq.retain_grad()
    t = so._ln(sd, p + ".norm2", t + so.mha(sd, p + ".multihead_attn", cq, memory, memory, None), so.LN_EPS_DETR)
    ff = so._lin(sd, p + ".linear2", F.relu(so._lin(sd, p + ".linear1", t)))
    cap.append((p, qk, cq))
    return so._ln(sd, p + ".norm3", t + ff, so.LN_EPS_DETR)
so.decoder_layer = dl
inp2 = dict(inp); inp2["visn_tokens"] = tokens
o = so.agqa_forward(W, inp2)
idx_r = [(torch.from_numpy(g["rel_rows"][f][g["rel_rows"][f] >= 0]), torch.from_numpy(g["rel_cols"][f][g["rel_cols"][f] >= 0])) for f in range(32)]
idx_a = [(torch.from_numpy(g["act_rows"][f][g["act_rows"][f] >= 0]), torch.from_numpy(g["act_cols"][f][g["act_cols"][f] >= 0])) for f in range(32)]
qa = torch.nn.functional.binary_cross_entropy_with_logits(o["hg_logit"], inp["target"]) * 171
l1, _, _ = so.set_loss(o["rel_logits"], inp["rel_labels"], idx_r)
l2, _, _ = so.set_loss(o["act_logits"], inp["act_labels"], idx_a)
o["rel_out"].retain_grad(); o["act_out"].retain_grad()
(qa + l1 + l2).backward()
for name, key, outk in (("rel", "relation_query_embed", "rel_out"), ("act", "action_query_embed", "act_out")):
    truth = leaves[key].grad.reshape(-1, 768)
    mine = dbg["dq_" + name[0]]
    print(name, "dq rel-l2 %.4f" % float((mine - truth).norm() / truth.norm()), "norms", float(mine.norm()), float(truth.norm()))
    print("  row0 mine %.4g truth %.4g  diff %.4g" % (float(mine[0].norm()), float(truth[0].norm()), float((mine[0]-truth[0]).norm())))
    per_row = ((mine - truth).norm(dim=1) / truth.norm(dim=1).clamp_min(1e-12))
    print("  per-row rel err: median %.4f max %.4f" % (float(per_row.median()), float(per_row.max())))
    t2 = o[outk].grad.reshape(-1, 768); m2 = dbg["d_%s_out" % name]
    print("  d_out rel-l2 %.4f" % float((m2 - t2).norm() / t2.norm()))
    # LN-backward sensitivities
    print("  colsum(mine) vs colsum(truth): %.4g vs %.4g ; diff %.4g" % (float(mine.sum(0).norm()), float(truth.sum(0).norm()), float((mine.sum(0)-truth.sum(0)).norm())))

tr_list = tr.engine._trace   # order: rel layers 4..0 (dt1p, dtgt_p each), then act layers 4..0
truth_terms = {}
for p, qk, cq in cap:
    truth_terms[p] = (qk.grad.reshape(-1, 768), cq.grad.reshape(-1, 768))
names = ["rel_decoder.layers.%d" % i for i in reversed(range(5))] + ["action_decoder.layers.%d" % i for i in reversed(range(5))]
k = 0
for n in names:
    tq, tc = truth_terms[n]
    (tag1, m_dt1p, _), (tag2, m_dtgt, _) = tr_list[k], tr_list[k + 1]
    k += 2
    print(n, tag1, "rel-l2 %.3f (|truth| %.3g)" % (float((m_dt1p - tc).norm() / tc.norm()), float(tc.norm())),
          tag2, "rel-l2 %.3f (|truth| %.3g |mine| %.3g)" % (float((m_dtgt - tq).norm() / tq.norm().clamp_min(1e-20)), float(tq.norm()), float(m_dtgt.norm())))

# kernel-only check: torch fp32 autograd on the SAME bf16 inputs of rel decoder layer 4 self-attention
qk, v, dsa, dqk, dv, sa = tr.engine._trace_self[0]
B, nq, per = 2, 128, 8
q_ = qk[:, :768].clone().requires_grad_(True); k_ = qk[:, 768:].clone().requires_grad_(True); v_ = v.clone().requires_grad_(True)
def heads(x): return x.view(B, nq, 12, 64).transpose(1, 2)
f = torch.arange(nq) // per
mask = torch.zeros(nq, nq); mask[f[None, :] > f[:, None]] = float("-inf")
sc = heads(q_) @ heads(k_).transpose(-1, -2) * 0.125 + mask
o_ = (torch.softmax(sc, -1) @ heads(v_)).transpose(1, 2).reshape(B * nq, 768)
o_.backward(dsa)
print("fwd sa vs torch: %.4f" % float((sa - o_.detach()).norm() / o_.detach().norm()))
for nm, mine, ref in (("dq", dqk[:, :768], q_.grad), ("dk", dqk[:, 768:], k_.grad), ("dv", dv, v_.grad)):
    print("self-attn L4 %s: rel-l2 %.4f  |ref| %.4g |mine| %.4g" % (nm, float((mine - ref).norm() / ref.norm()), float(ref.norm()), float(mine.norm())))
vh = heads(v)
print("V row variation (per head): |V - mean| / |V| = %.4g" % float((vh - vh.mean(2, keepdim=True)).norm() / vh.norm()))
