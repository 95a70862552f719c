"""Host-side mirror of the reference's Python interface for the hot path.

Same names, constructor arguments, call signatures, return structures and error behaviour as the symbols the
reference's task scripts import, routed through libshgvqa:

  * ``HungarianMatcher``  ≙ lxrt/matcher.py:14-104
  * ``AGQAHotPath``       ≙ tasks/agqa_model.py:16-269 (``AGQAModel.forward``'s 5-tuple; ``state_dict`` layout)

The slow_r50 backbone and the Conv3d front-end are NOT part of the path (north_star): ``AGQAHotPath`` runs the
front-end with stock PyTorch ops on the reference's own weights and hands [B,393,768] tokens to the engine.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .engine import Engine
from .params import ParamStore, build_key_shapes


class HungarianMatcher(nn.Module):
    """Drop-in for lxrt/matcher.py's HungarianMatcher: softmax → cost → per-frame LSAP, on the device."""

    def __init__(self, cost_class: float = 1, loss_hg_per_frame: bool = False, clip_len: int = 16):
        super().__init__()
        self.cost_class = cost_class
        self.loss_hg_per_frame = loss_hg_per_frame
        self.clip_len = clip_len
        assert cost_class != 0, "cost cant be 0"               # matcher.py:36
        if cost_class < 0:
            raise NotImplementedError("negative cost_class flips the optimum; the reference never uses it")

    @torch.no_grad()
    def forward(self, outputs, targets):
        logits = outputs["pred_logits"]
        bs, num_queries, n_cls = logits.shape
        if self.loss_hg_per_frame:
            flat = [t for sub in targets for t in sub["labels"]]          # matcher.py:65
            n_frames, q_f = bs * self.clip_len, num_queries // self.clip_len
        else:
            flat = [v["labels"] for v in targets]                         # matcher.py:82
            n_frames, q_f = bs, num_queries
        if q_f > 32 or any(int(t.numel()) > 32 for t in flat):
            raise NotImplementedError("device LSAP handles up to 32 queries x 32 targets per problem")
        sizes = [int(t.numel()) for t in flat]
        if sum(sizes) == 0:
            # the reference's `.view(..., -1)` on a 0-element tensor raises here (matcher.py:76)
            raise RuntimeError("cannot reshape tensor of 0 elements into shape [%d, %d, -1]" % (n_frames, q_f))
        dev = logits.device
        labels = torch.cat([t.reshape(-1) for t in flat]).to(device=dev, dtype=torch.int32)
        off = torch.zeros(n_frames + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(torch.tensor(sizes, dtype=torch.int64), 0).to(torch.int32)
        lg = logits.reshape(bs * num_queries, n_cls)
        if lg.dtype not in (torch.float32, torch.bfloat16):
            lg = lg.float()
        m = ops.match_from_logits(lg.contiguous(), labels, off.to(dev), q_f, want_targets=False)
        rows, cols, n, status = (m[k].cpu() for k in ("rows", "cols", "n_match", "status"))   # one sync
        if bool((status != 0).any()):
            bad = int(status[status != 0][0])
            raise ValueError("cost matrix is infeasible" if bad == -1 else "matrix contains invalid numeric entries")
        rows, cols = rows.to(torch.int64), cols.to(torch.int64)
        return [(rows[f, :int(n[f])].clone(), cols[f, :int(n[f])].clone()) for f in range(n_frames)]


def build_matcher(args):
    return HungarianMatcher(cost_class=args.set_cost_class)              # matcher.py:107-108


class AGQAHotPath(nn.Module):
    """Inference-side mirror of AGQAModel: same forward signature and 5-tuple, same state_dict keys."""

    def __init__(self, num_answers, num_queries=128, num_classes=456, num_actions=156, task="hgqa", device="cuda"):
        super().__init__()
        self.task = task
        self.key_shapes = build_key_shapes(task, num_answers, num_classes + 1, num_actions + 1, n_rel_q=num_queries)
        self.store = ParamStore(self.key_shapes, device, with_grad=False)
        self.engine = Engine(self.store, task)
        self.max_seq_length = 40

    @classmethod
    def from_reference_module(cls, model, task="hgqa"):
        sd = {k: v for k, v in model.state_dict().items() if not k.startswith("vid_encoder.")}
        n_ans = sd["logit_fc.3.weight"].shape[0]
        n_rel = sd["class_embed.3.weight"].shape[0] - 1 if task == "hgqa" else 456
        n_act = sd["action_embed.3.weight"].shape[0] - 1 if task == "hgqa" else 156
        nq = sd["relation_query_embed.word_embeddings.weight"].shape[0] if task == "hgqa" else 128
        self = cls(n_ans, nq, n_rel, n_act, task)
        self.load_state_dict(sd, strict=True)
        return self

    def state_dict(self, *a, **k):
        return self.store.state_dict()

    def load_state_dict(self, sd, strict=True):
        r = self.store.load_state_dict(sd, strict)
        self.engine.refresh_derived()
        return r

    @torch.no_grad()
    def forward(self, feat, pos, input_ids, input_masks, segment_ids, rel_segment_ids=None, rel_tgt_mask=None,
                act_segment_ids=None, act_tgt_mask=None, hg_mask=None, rel_tgt_ids=None, act_tgt_ids=None):
        inp = dict(feat=feat.float().contiguous(), input_ids=input_ids, input_mask=input_masks,
                   segment_ids=segment_ids, rel_segment_ids=rel_segment_ids, act_segment_ids=act_segment_ids)
        out, _ = self.engine.forward(inp, train=False)
        b = input_ids.shape[0]
        attn_probs = ([], [], [], [], [])                                  # empty lists, as under --noCaps (:1850-1854)
        if self.task != "hgqa":
            return out["logit"], attn_probs
        nq = rel_segment_ids.shape[1]
        return (out["logit"], out["rel_logits"].reshape(b, nq, -1), out["act_logits"].reshape(b, -1, out["act_logits"].shape[-1]),
                out["hg_logit"], attn_probs)
