"""Host-side mirror of the reference's Python interface for the hot path.

Same names, constructor arguments, call signatures, return structures and error behaviour as the symbols the
reference's task scripts import, routed through libshgvqa:

  * ``HungarianMatcher``  ≙ lxrt/matcher.py:14-104
  * ``AGQAHotPath``       ≙ tasks/agqa_model.py:16-269 (``AGQAModel``: forward 5-tuple / 2-tuple, ``state_dict`` layout,
                            ``parameters()`` with ``.grad`` — the reference's own train loop, tasks/agqaHGQA.py:326-392,
                            runs on it unchanged: its matcher, ``loss_labels``, ``total_loss.backward()``,
                            ``clip_grad_norm_`` and ``BertAdam``)
  * ``LXRTEncoder``       ≙ lxrt/entry.py:139-201 (``(feats, pooled, attn_probs)``, ``attn_probs[-1]`` the 4-tuple of
                            lxrt/modeling_capsbert.py:1297-1302)

How training through the boundary works.  The engine (engine.py) is a hand-written forward/backward over flat
parameter buffers; ``_HotPathFn`` wraps it in a ``torch.autograd.Function`` whose outputs are the logits.  Every
parameter is exposed as an ``nn.Parameter`` that is a VIEW of the flat fp32 master buffer, and after a backward its
``.grad`` is a VIEW of the flat fp32 gradient buffer — so ``optimizer.zero_grad()``, ``clip_grad_norm_`` and any
torch optimiser (the reference's BertAdam included) act directly on the engine's storage.  The bf16 shadow the
tensor cores read is refreshed from the master at the next forward.

The slow_r50 backbone is NOT part of the path (north_star): pass it as ``vid_encoder`` (any module with
``encode(clip) -> [B,2048,T,7,7]``) and it runs as stock PyTorch, its gradient arriving through autograd; without
one, ``feat`` must already be the backbone's output.  The Conv3d front-end runs inside the engine (csrc/conv.cu).
"""
import collections

import torch
from torch import nn

from . import ops
from .engine import Engine
from .params import ParamStore, build_key_shapes, canonical_key

BF16 = torch.bfloat16


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
        lg = logits.detach().reshape(bs * num_queries, n_cls)
        if lg.dtype not in (torch.float32, torch.bfloat16):
            lg = lg.float()
        m = ops.match_from_logits(lg.contiguous(), labels, off.to(dev), q_f, want_targets=False)
        rows, cols, n, status = (m[k].cpu() for k in ("rows", "cols", "n_match", "status"))   # one sync
        if bool((status != 0).any()):
            bad = int(status[status != 0][0])
            if bad == -4:
                raise IndexError("target label out of range for %d classes" % n_cls)      # prob[:, tgt_ids], matcher.py:73
            raise ValueError("cost matrix is infeasible" if bad == -1 else "matrix contains invalid numeric entries")
        rows, cols = rows.to(torch.int64), cols.to(torch.int64)
        return [(rows[f, :int(n[f])].clone(), cols[f, :int(n[f])].clone()) for f in range(n_frames)]


def build_matcher(args):
    return HungarianMatcher(cost_class=args.set_cost_class)              # matcher.py:107-108


def _pad_bf16(d, rows, n):
    """fp32/bf16 gradient [.., n] -> bf16 [rows, n] view with a leading dimension that is a multiple of 8."""
    ld = (n + 7) // 8 * 8
    buf = torch.zeros(rows, ld, device=d.device, dtype=BF16)
    buf[:, :n].copy_(d.reshape(rows, n))
    return buf[:, :n]


class _HotPathFn(torch.autograd.Function):
    """The engine's forward/backward as one autograd node.  Inputs: the owning module, the backbone features and an
    anchor that requires grad (so that the node is always recorded); parameter gradients do not travel through
    autograd — they land in the flat gradient buffer, which the parameters' ``.grad`` views alias."""

    @staticmethod
    def forward(ctx, owner, anchor, feat, inp):
        out, ectx = owner.engine.forward(inp, train=True)
        ctx.owner, ctx.ectx = owner, ectx
        ctx.set_materialize_grads(False)
        b = inp["input_ids"].shape[0]
        ctx.b = b
        res = [out["logit"].float().reshape(b, -1)]
        if owner.task == "hgqa":
            res += [out["rel_logits"].float().reshape(b, -1, out["rel_logits"].shape[-1]),
                    out["act_logits"].float().reshape(b, -1, out["act_logits"].shape[-1]),
                    out["hg_logit"].float().reshape(b, -1)]
        ctx.shapes = [tuple(r.shape) for r in res]
        owner._last_out = out
        return tuple(res)

    @staticmethod
    def backward(ctx, *grads):
        owner, ectx, b = ctx.owner, ctx.ectx, ctx.b
        eng = owner.engine
        owner._begin_backward()

        def prep(g, shape):
            rows, n = shape[0] * (shape[1] if len(shape) == 3 else 1), shape[-1]
            if g is None:
                g = torch.zeros(shape, device=eng.dev, dtype=torch.float32)
            return _pad_bf16(g, rows, n)
        d_logit = prep(grads[0], ctx.shapes[0]) if grads[0] is not None else None
        need_dfeat = ctx.needs_input_grad[2]
        if owner.task == "hgqa":
            d_rel, d_act, d_hg = (prep(grads[i], ctx.shapes[i]) for i in (1, 2, 3))
            eng.backward(ectx, d_hg, d_rel, d_act, need_dvisn=need_dfeat, d_logit=d_logit)
        else:
            if d_logit is None:
                d_logit = prep(None, ctx.shapes[0])
            eng.backward(ectx, None, None, None, need_dvisn=need_dfeat, d_logit=d_logit)
        owner._end_backward()
        d_feat = ectx.get("d_feat") if need_dfeat else None
        return None, None, d_feat, None


class FeatureBackbone(nn.Module):
    """Stand-in for video_encoder.py's VideoBackbone when the caller feeds backbone FEATURES: encode is the
    identity, and `model.module.vid_encoder.eval()` (tasks/agqaHGQA.py:263) has something to call."""

    def encode(self, x):
        return x

    def forward(self, x):
        return x


class _Node(nn.Module):
    """Container of the parameter tree that mirrors the reference's module names (see build_param_tree)."""


class _LxrtNode(_Node):
    """What tasks/agqaHGQA.py touches on ``model.lxrt_encoder`` (lxrt/entry.py:139-238): dim, load, multi_gpu."""

    dim = 768

    def multi_gpu(self):
        pass       # one process per GPU: there is no nn.DataParallel replica to build

    def load(self, path):
        """entry.py:214-238: `<path>_LXRT.pth`, keys optionally prefixed `module.`, loaded non-strictly into
        `self.model`."""
        sd = torch.load("%s_LXRT.pth" % path, map_location="cpu")
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        own = self._modules["model"].state_dict()
        self._modules["model"].load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)


def build_param_tree(root, store, key_shapes):
    """Registers every tensor of the reference's state_dict on `root` under the reference's own dotted path
    (nested container modules), each nn.Parameter a VIEW of the store's flat fp32 master buffer.  Aliased keys
    (the shared CrossLayer under cross_attn_layer.cross / x_layers.0 / x_layers.1, pooler = pooler_dict.cross;
    lxrt/modeling_capsbert.py:1122-1124,1247-1249,1144) get the SAME Parameter object, like the reference's shared
    modules — so the stock nn.Module machinery (parameters(), state_dict(), load_state_dict(strict=True),
    nn.DataParallel) sees exactly the reference's 669 / 324 keys and its de-duplicated parameter list.
    Returns {canonical key: Parameter}."""
    params = collections.OrderedDict()
    for k, (shape, dtype) in key_shapes.items():
        parts = k.split(".")
        mod = root
        for name in parts[:-1]:
            if name not in mod._modules:
                mod.add_module(name, _LxrtNode() if (mod is root and name == "lxrt_encoder") else _Node())
            mod = mod._modules[name]
        if dtype != "float32":
            mod.register_buffer(parts[-1], store.int_buffers[k])
            continue
        ck = canonical_key(k)
        if ck not in params:
            params[ck] = nn.Parameter(store.p(ck), requires_grad=True)
        mod.register_parameter(parts[-1], params[ck])
    return params


class AGQAHotPath(nn.Module):
    """Mirror of tasks/agqa_model.py's AGQAModel over the hand-written engine: same forward signature and returns
    (5-tuple for --taskHGQA, (logit, attn_probs) for --taskVQA), same state_dict keys, trainable through
    ``loss.backward()`` with any torch optimiser over ``parameters()``.

    vid_encoder : module with ``encode(clip)`` (the stock slow_r50 wrapper, video_encoder.py:30-51) or None — then
                  ``feat`` is the backbone's output [B,2048,T,7,7]; a 3-channel clip is refused with a clear error.
    dropout     : training-mode dropout at the reference's probabilities when ``self.training``.
    use_hg_mask : --useHGMask (tasks/agqa_model.py:261-264): ``hg_mask`` [B,T,11] masks hypergraph keys.
    vqa_head_grad: also backpropagate through the LXRT answer logits `logit` in --taskHGQA mode (off by default:
                  the reference's HGQA loss never reads it, and the engine then runs that branch forward-only).
    """

    def __init__(self, num_answers, num_queries=128, num_classes=456, num_actions=156, task="hgqa", device="cuda",
                 vid_encoder=None, dropout=True, use_hg_mask=False, vqa_head_grad=False, dropout_seed=0):
        super().__init__()
        self.task = task
        self.key_shapes = build_key_shapes(task, num_answers, num_classes + 1, num_actions + 1, n_rel_q=num_queries)
        self.store = ParamStore(self.key_shapes, device, with_grad=True)
        self.engine = Engine(self.store, task)
        self.engine.drop_seed.fill_(dropout_seed & 0x7fffffff)
        self.max_seq_length = 40                                           # tasks/agqa_model.py:14
        self.dropout, self.use_hg_mask, self.vqa_head_grad = bool(dropout), bool(use_hg_mask), bool(vqa_head_grad)
        self._anchor = torch.zeros((), device=self.store.device, requires_grad=True)
        self._params = build_param_tree(self, self.store, self.key_shapes)
        self.vid_encoder = vid_encoder if vid_encoder is not None else FeatureBackbone()
        self._dirty = True
        self._last_out = None
        self.register_load_state_dict_post_hook(AGQAHotPath._mark_dirty)

    @staticmethod
    def _mark_dirty(module, incompatible_keys):
        module._dirty = True

    # ------------------------------------------------------------------ nn.Module surface
    def _apply(self, fn, recurse=True):
        # .to(device) / .cuda() / .float(): the flat buffers already live on the B200 and stay fp32 there; only a
        # stock backbone is moved
        self.vid_encoder._apply(fn)
        return self

    @classmethod
    def from_reference_module(cls, model, task="hgqa", **kw):
        sd = {k: v for k, v in model.state_dict().items() if not k.startswith("vid_encoder.")}
        n_ans = sd["logit_fc.3.weight"].shape[0]
        n_rel = sd["class_embed.3.weight"].shape[0] - 1 if task == "hgqa" else 456
        n_act = sd["action_embed.3.weight"].shape[0] - 1 if task == "hgqa" else 156
        nq = sd["relation_query_embed.word_embeddings.weight"].shape[0] if task == "hgqa" else 128
        self = cls(n_ans, nq, n_rel, n_act, task, **kw)
        self.load_state_dict(sd, strict=True)
        return self

    def load_state_dict(self, sd, strict=True, assign=False):
        """Reference semantics (tasks/agqaHGQA.py:863-870): optional `module.` prefix, strict key check.  A
        checkpoint without backbone keys loads into a model fed with backbone features."""
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        if isinstance(self.vid_encoder, FeatureBackbone):
            sd = {k: v for k, v in sd.items() if not k.startswith("vid_encoder.")}
        r = super().load_state_dict(sd, strict=strict)
        self._dirty = True
        return r

    def sync_weights(self):
        """bf16 shadow / derived operands <- fp32 master (after an optimiser or a load wrote the parameters)."""
        self.store.refresh_shadow()
        self.engine.refresh_derived()
        self._dirty = False

    # ------------------------------------------------------------------ gradient plumbing
    def _begin_backward(self):
        """torch semantics: a backward ACCUMULATES into .grad when it exists and creates it otherwise.  The engine
        accumulates into the flat buffer, so views whose .grad is None (optimizer.zero_grad(set_to_none=True), the
        reference's call at tasks/agqaHGQA.py:264) are zeroed first."""
        fresh = [k for k, p in self._params.items() if p.grad is None]
        if len(fresh) == len(self._params):
            self.store.grad.zero_()
        else:
            for k in fresh:
                self.store.g(k).zero_()
        self._fresh = fresh

    def _end_backward(self):
        """Attach the gradient views.  Tensors the loss does not reach keep .grad = None, as in the reference
        (its BertAdam skips them, lxrt/optimization.py:122-123): the dead ModuleDict branches always, the LXRT
        cross layers / pooler in --taskHGQA unless vqa_head_grad."""
        for k in self._fresh:
            if self._reached(k):
                self._params[k].grad = self.store.g(k)
        self._dirty = True

    def _reached(self, k):
        from .trainer import live_with_frontend
        if ".visn_fc." in k:
            return live_with_frontend(k)
        dead = (".cross_attn_layer.self.", ".cross_attn_layer.cross_self.", ".cross_attn_layer.old.",
                "pooler_dict.self.", "pooler_dict.cross_self.", "pooler_dict.no_cross.", "pooler_dict.old.")
        if any(d in k for d in dead):
            return False
        if k.startswith("lxrt_encoder.") and (".cross_attn_layer.cross." in k or ".pooler_dict.cross." in k):
            return self.task != "hgqa" or self.vqa_head_grad
        return True

    # ------------------------------------------------------------------ forward
    def _inputs(self, feat, input_ids, input_masks, segment_ids, rel_segment_ids, act_segment_ids, hg_mask):
        feat = self.vid_encoder.encode(feat)                              # tasks/agqa_model.py:197
        if feat.dim() != 5 or feat.shape[1] != 2048 or feat.shape[-2:] != (7, 7):
            raise ValueError("AGQAHotPath: `feat` must be the backbone output [B,2048,T,7,7] (got %s); pass the "
                             "backbone as vid_encoder= to feed raw clips" % (tuple(feat.shape),))
        inp = dict(feat=feat.detach().float().contiguous(), input_ids=input_ids, input_mask=input_masks,
                   segment_ids=segment_ids)
        if self.task == "hgqa":
            inp["rel_segment_ids"], inp["act_segment_ids"] = rel_segment_ids, act_segment_ids
            if self.use_hg_mask and hg_mask is not None:
                # CrossEncoder.forward, lxrt/modeling_capsbert.py:1187-1194: [cls = 1 | hg_mask] -> (1 - m) * -10000
                b = input_ids.shape[0]
                m = torch.cat([torch.ones(b, 1, device=hg_mask.device), hg_mask.reshape(b, -1).float()], dim=1)
                inp["hg_key_mask"] = ((1.0 - m) * -10000.0).contiguous()
            inp["x_grad"] = self.vqa_head_grad
        return feat, inp

    def forward(self, feat, pos, input_ids, input_masks, segment_ids, rel_segment_ids=None, rel_tgt_mask=None,
                act_segment_ids=None, act_tgt_mask=None, hg_mask=None, rel_tgt_ids=None, act_tgt_ids=None):
        """tasks/agqa_model.py:166-269.  `pos` (boxes) is all ones in the reference's data path and only becomes an
        all-zero additive mask (lxrt/modeling_capsbert.py:1836); rel_tgt_mask / act_tgt_mask are ignored there too and
        regenerated (agqa_model.py:220,241-242) — the block-causal frame mask is applied inside the attention kernel."""
        if rel_tgt_ids is not None or act_tgt_ids is not None:
            raise NotImplementedError("--gtHG (ground-truth hypergraph inputs) is outside the README's configuration")
        feat, inp = self._inputs(feat, input_ids, input_masks, segment_ids, rel_segment_ids, act_segment_ids, hg_mask)
        if self._dirty:
            self.sync_weights()
        attn_probs = ([], [], [], [], [])                                  # empty lists, as under --noCaps (:1850-1854)
        grad_mode = self.training and torch.is_grad_enabled()
        self.engine.dropout_on = self.dropout and self.training
        if grad_mode:
            if self.engine.dropout_on:
                self.engine.drop_seed.add_(1)
            res = _HotPathFn.apply(self, self._anchor, feat, inp)
        else:
            with torch.no_grad():
                self.engine._training = False
                out, _ = self.engine.forward(inp, train=False)
            self._last_out = out
            b = input_ids.shape[0]
            res = [out["logit"].float().reshape(b, -1)]
            if self.task == "hgqa":
                res += [out["rel_logits"].float().reshape(b, -1, out["rel_logits"].shape[-1]),
                        out["act_logits"].float().reshape(b, -1, out["act_logits"].shape[-1]),
                        out["hg_logit"].float().reshape(b, -1)]
        if self.task != "hgqa":
            return res[0], attn_probs
        return res[0], res[1], res[2], res[3], attn_probs


class LXRTEncoder(nn.Module):
    """Mirror of lxrt/entry.py's LXRTEncoder over the engine (--noCaps, crossAttnType cross): language, relation and
    cross layers + BertPooler2, inference.  forward((ids, mask, seg), (feat, boxes)) -> (feats, pooled, attn_probs):
    mode 'lxr' returns feats = (lang [B,40,768], visn [B,393,768]), mode 'x' returns feats = None (entry.py:187-201);
    attn_probs[-1] = (language-layer output, additive language mask [B,1,1,40], relation-layer output, visual mask)
    as lxrt/modeling_capsbert.py:1297-1302 returns it and tasks/agqa_model.py:218 consumes it."""

    def __init__(self, args=None, max_seq_length=40, mode="x", device="cuda", num_answers=171):
        super().__init__()
        self.max_seq_length, self.mode, self.args = max_seq_length, mode, args
        self.key_shapes = build_key_shapes("vqa", num_answers)
        self.store = ParamStore(self.key_shapes, device, with_grad=False)
        self.engine = Engine(self.store, "vqa")

    @property
    def dim(self):
        return 768

    def load_state_dict(self, sd, strict=True):
        """Takes the full model's keys (`lxrt_encoder.model.bert...`) or the encoder's own (`model.bert...`)."""
        if not any(k.startswith("lxrt_encoder.") for k in sd):
            sd = {"lxrt_encoder." + k: v for k, v in sd.items()}
        sd = {k: v for k, v in sd.items() if k in self.key_shapes}
        missing = [k for k in self.key_shapes if k.startswith("lxrt_encoder.") and k not in sd]
        if strict and missing:
            raise RuntimeError("state_dict mismatch: missing %s" % missing[:5])
        return self.store.load_state_dict(sd, strict=False)

    def state_dict(self, *a, **k):
        return collections.OrderedDict((kk[len("lxrt_encoder."):], v) for kk, v in self.store.state_dict().items()
                                       if kk.startswith("lxrt_encoder."))

    @torch.no_grad()
    def forward(self, sents, feats, visual_attention_mask=None):
        input_ids, input_mask, segment_ids = sents[0], sents[1], sents[2]
        feat = feats[0] if isinstance(feats, (tuple, list)) else feats
        inp = dict(feat=feat.float().contiguous(), input_ids=input_ids, input_mask=input_mask,
                   segment_ids=segment_ids)
        out, _ = self.engine.forward(inp, train=False)
        b, nl = input_ids.shape
        nv = out["visn_r"].shape[0] // b
        lang_mask = out["kmask"].view(b, 1, 1, nl)
        visn_mask = torch.zeros(b, 1, 1, nv, device=lang_mask.device)
        extra = (out["lang_l"].float().view(b, nl, 768), lang_mask, out["visn_r"].float().view(b, nv, 768), visn_mask)
        attn_probs = ([], [], [], [], [], extra)
        pooled = out["pooled"].float()
        f = (out["lang"].float().view(b, nl, 768), out["visn"].float().view(b, nv, 768)) if self.mode == "lxr" else None
        return f, pooled, attn_probs
