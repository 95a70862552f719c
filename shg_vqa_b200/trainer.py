"""One data-parallel training / evaluation step of the hot path.

Mirrors what tasks/agqaHGQA.py:326-392 does around the model call — answer loss,
per-frame Hungarian matching for relations and actions, weighted CE, backward,
clip_grad_norm_(5.), BertAdam — but keeps every stage on the device and on one
stream: ragged per-frame targets travel as CSR (labels + offsets) instead of
2*16*B tiny tensors, the matcher never does `.cpu()`, and the only collective is an
NCCL all-reduce of the flat gradient buffer (one process per GPU, replacing
nn.DataParallel's per-step scatter/replicate/gather).
"""
import math
import os

import torch

from . import ops
from .engine import Engine
from .params import ParamStore, build_key_shapes, canonical_key

BF16 = torch.bfloat16


def live_in_hgqa_training(key):
    """Parameters that receive a gradient in --taskHGQA training (the LXRT x-layers / pooler and
    the dead ModuleDict branches never do; tests/golden/hgqa_b2.npz `nograd_names` pins the set).  The Conv3d
    front-end is live in the reference but lies outside this path (stock PyTorch) and is excluded here."""
    k = canonical_key(key)
    dead = (".cross_attn_layer.self.", ".cross_attn_layer.cross_self.", ".cross_attn_layer.old.",
            "pooler_dict.self.", "pooler_dict.cross_self.", "pooler_dict.no_cross.", "pooler_dict.old.",
            ".visn_fc.")
    if any(d in k for d in dead):
        return False
    if k.startswith("lxrt_encoder.") and (".cross_attn_layer.cross." in k or ".pooler_dict.cross." in k):
        return False
    return True


def live_with_frontend(key):
    """live_in_hgqa_training plus the Conv3d front-end's own parameters (conv weights/biases, cls token, learned
    positions — live in the reference; box_fc / pos_layer_norm never receive a gradient), for runs that start from
    backbone features (`feat`) instead of precomputed visual tokens."""
    k = canonical_key(key)
    if ".visn_fc." in k:
        return ".visn_fc.conv." in k or k.endswith((".visn_fc.cls_token", ".visn_fc.position_encoding.pe.weight"))
    return live_in_hgqa_training(key)


def live_in_vqa_training(key):
    """Parameters that receive a gradient in --taskVQA training (tasks/agqaVQA.py: BCE on `logit` only): embeddings,
    language / relation / cross layers, the LXRT pooler and logit_fc, plus the Conv3d front-end's own parameters."""
    k = canonical_key(key)
    dead = (".cross_attn_layer.self.", ".cross_attn_layer.cross_self.", ".cross_attn_layer.old.",
            "pooler_dict.self.", "pooler_dict.cross_self.", "pooler_dict.no_cross.", "pooler_dict.old.")
    if any(d in k for d in dead):
        return False
    if ".visn_fc." in k:
        return ".visn_fc.conv." in k or k.endswith((".visn_fc.cls_token", ".visn_fc.position_encoding.pe.weight"))
    return True


def warmup_linear(x, warmup=0.002):
    """lxrt/optimization.py:38-43."""
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0.0)


def csr_targets(frame_labels, device):
    """List of B*16 ragged int64 label vectors (matcher.py:65 order) → (labels int32, offsets int32) on device."""
    sizes = torch.tensor([0] + [int(l.numel()) for l in frame_labels], dtype=torch.int64)
    off = torch.cumsum(sizes, 0).to(torch.int32)
    flat = torch.cat([l.reshape(-1) for l in frame_labels]).to(torch.int32) if int(off[-1]) else \
        torch.zeros(0, dtype=torch.int32)
    return flat.to(device), off.to(device)


def csr_from_padded(tokens, lengths):
    """Vectorised form of the reference's per-sample Python featuriser (lxrt/entry.py:76-97): padded
    [B,16,q_f] labels + [B,16] lengths → CSR, on whatever device the inputs live."""
    b, t, q = tokens.shape
    keep = torch.arange(q, device=tokens.device)[None, None, :] < lengths[:, :, None]
    flat = tokens[keep].to(torch.int32)
    off = torch.zeros(b * t + 1, dtype=torch.int32, device=tokens.device)
    off[1:] = torch.cumsum(lengths.reshape(-1), 0).to(torch.int32)
    return flat, off


def all_reduce_live_grads(store, world_size, group=None):
    """The path's only exchange step: SUM the contiguous live-gradient span of the flat buffer
    over the data-parallel ranks (NCCL over NVLink on GPUs; gloo in the CPU tests).  Averaging is left to the
    optimiser's grad_scale = 1/world_size, so no extra pass over the buffer is needed."""
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(store.grad[store.live_extent], op=dist.ReduceOp.SUM, group=group)


def shard_seed(base_seed, rank):
    """Each rank draws its own 32 clips (weak scaling): disjoint synthetic shards by seed."""
    return base_seed + rank


class HotPathTrainer:
    def __init__(self, device, task="hgqa", n_answers=171, n_rel_cls=457, n_act_cls=158, lr=1e-5, t_total=-1,
                 warmup=0.1, weight_decay=0.01, max_grad_norm=5.0, eos_coef=0.1, world_size=1, process_group=None,
                 live=None, dropout=False, dropout_seed=0, rank=0, frontend=False,
                 backbone_grad=False, global_ce_norm=False):
        self.device = torch.device(device)
        # frontend: the step starts from backbone features (inp['feat']) and trains the Conv3d front-end too;
        # backbone_grad: also return the gradient wrt those features (ctx['d_feat']) for a trainable backbone
        self.frontend, self.backbone_grad = bool(frontend), bool(backbone_grad)
        if live is None:
            if task != "hgqa":
                live = live_in_vqa_training if frontend else (lambda k: live_in_vqa_training(k) and ".visn_fc." not in k)
            else:
                live = live_with_frontend if frontend else live_in_hgqa_training
        self.task = task
        self.key_shapes = build_key_shapes(task, n_answers, n_rel_cls, n_act_cls)
        self.store = ParamStore(self.key_shapes, self.device, live=live)
        self.engine = Engine(self.store, task)
        self.lr, self.t_total, self.warmup = lr, t_total, warmup
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-6           # BertAdam defaults (lxrt/optimization.py:63-64)
        self.wd, self.max_norm, self.eos = weight_decay, max_grad_norm, eos_coef
        self.world_size, self.pg = world_size, process_group
        self.step_count = 0
        self.exp_avg = self.exp_avg_sq = None
        self.n_ans = n_answers
        self.scal = torch.zeros(16, device=self.device, dtype=torch.float32)   # loss sums / grad sumsq / error count
        self.scal_den = torch.zeros(8, device=self.device, dtype=torch.float32)   # globally reduced CE sums
        self.sumsq_ws = torch.zeros(2048, device=self.device, dtype=torch.float32)   # shg_grad_sumsq's partials + counter
        self.global_ce_norm = bool(global_ce_norm)
        # persistent-GEMM grid while the step's all-reduce is in flight (0 = all 148 SMs); SHG_DP_GEMM_CTAS overrides
        self.dp_gemm_ctas = int(os.environ.get("SHG_DP_GEMM_CTAS", "0"))
        self.lr_dev = torch.zeros(1, device=self.device, dtype=torch.float32)   # device-resident lr (graph-safe)
        # lr crosses to the device through a ring of pinned slots, one event each: the host runs ahead of the GPU
        # (lagged loss read), and a single slot could be overwritten before its queued H2D copy has read it
        self._lr_ring = 8
        cuda = self.device.type == "cuda"
        self.lr_host = torch.zeros(self._lr_ring, dtype=torch.float32).pin_memory() if cuda else None
        self._lr_events = [None] * self._lr_ring
        self._lr_slot = 0
        self.graph = None
        self.copy_stream, self._staged = None, False
        self._loss_slots = None
        # training-mode dropout (the reference trains with model.train(): hidden/attention 0.1, decoders 0.15).
        # Masks are counter-based functions of (step seed, call site, element); the seed lives on the device and
        # is advanced at the start of every step (inside the captured graph too), and differs per rank.
        self.engine.dropout_on = bool(dropout)
        if self.engine.drop_seed is not None:
            self.engine.drop_seed.fill_((dropout_seed + 1000003 * rank) & 0x7fffffff)

    # ------------------------------------------------------------------ weights
    def init_weights(self, seed=9595):
        """Random init with the reference's distributions (--fromScratch: init_bert_weights,
        lxrt/modeling_capsbert.py:1640-1651 — N(0, 0.02) matrices/embeddings, LayerNorm 1/0, zero biases;
        HGEmbeddings keep nn.Embedding's N(0,1) with a zero padding row; zero tokens)."""
        st = self.store
        g = torch.Generator(device=self.device).manual_seed(seed)
        st.master.zero_()
        for k, (off, n, shape) in st.offsets.items():
            v = st.master[off:off + n].view(shape)
            if k.endswith(("LayerNorm.weight", "norm1.weight", "norm2.weight", "norm3.weight", "layer_norm.weight")) \
                    or k in ("class_embed.2.weight", "action_embed.2.weight", "logit_fc.2.weight"):
                v.fill_(1.0)
            elif k.startswith(("relation_query_embed.", "action_query_embed.")) and k.endswith("embeddings.weight"):
                v.normal_(0.0, 1.0, generator=g)
                v[0].zero_()
            elif ".visn_fc.conv." in k:
                # nn.Conv3d keeps torch's default init (init_bert_weights only touches Linear/Embedding/LayerNorm):
                # U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias, fan_in = c_in * 45
                c_in = 2048 if ".conv.1." in k else 768
                bound = 1.0 / math.sqrt(c_in * 45)
                v.uniform_(-bound, bound, generator=g)
            elif len(shape) >= 2 and not k.endswith(("cls_token", "act_token", "rel_token")):
                v.normal_(0.0, 0.02, generator=g)
        st.refresh_shadow()
        self.engine.refresh_derived()

    def load_state_dict(self, sd, strict=True):
        r = self.store.load_state_dict(sd, strict)
        self.engine.refresh_derived()
        return r

    def state_dict(self):
        return self.store.state_dict()

    # ------------------------------------------------------------------ loss on device
    def loss_and_grads(self, out, batch, want_grad=True, ce_sums=None):
        """batch: target fp32 [B, n_ans]; rel_labels/rel_off, act_labels/act_off int32 CSR.  Optional
        rel_target_classes / act_target_classes (int32 [B*Q]) replace the matcher's scatter — used by the parity
        tests to hold the reference's assignment fixed.  ce_sums (fp32 [8] on the device: the relation and the action
        (sum w*nll, sum w, matched, correct)) replaces this batch's own sums as the weighted-CE normaliser — the
        data-parallel check uses it to give the single-process run the very denominators the ranks reduced (a
        last-bit difference there is amplified by the bf16 roundings of the backward).
        Returns (scalars tensor view, d_hg_logit, d_rel_logits, d_act_logits, match dicts).
        scal[0]=qa loss, [1..4]=rel sums (w*nll, w, matched, correct), [5..8]=act sums, [9]=grad sumsq,
        [10]=number of frames / rows the matcher or the loss refused (NaN/-inf costs, > 32 targets, label out of
        range — inputs the reference raises on); losses_from_scalars turns a non-zero count into a ValueError."""
        scal = self.scal
        scal.zero_()
        if self.task != "hgqa":
            # --taskVQA: BCEWithLogits(logit, target) * n_answers only (tasks/agqaVQA.py)
            d_logit = ops.bce_logits(out["logit"], batch["target"], scal[0:1], 1.0, want_grad)
            return scal, d_logit, None, None, []
        d_hg = ops.bce_logits(out["hg_logit"], batch["target"], scal[0:1], 1.0, want_grad)
        res = [d_hg]
        matches = []
        for fam, q_f, o in (("rel", self.engine.n_rel, 1), ("act", self.engine.n_act, 5)):
            lg = out[fam + "_logits"]
            m = ops.match_from_logits(lg, batch[fam + "_labels"], batch[fam + "_off"], q_f)
            if batch.get(fam + "_target_classes") is not None:
                m = dict(m, target_classes=batch[fam + "_target_classes"])
            ops.set_loss_fwd(lg, m, self.eos, out=scal[o:o + 4], err_out=scal[10:11])
            matches.append(m)
        gscale = 1.0
        if self.global_ce_norm and self.world_size > 1:
            # the reference's nn.DataParallel gathers the logits of all replicas and normalises the weighted CE by
            # the GLOBAL sum of weights (tasks/agqaHGQA.py:127,223): all-reduce the two denominators (8 bytes) and
            # undo the optimiser's 1/world average for these two losses
            import torch.distributed as dist
            den = self.scal_den
            den[0:4].copy_(scal[1:5])
            den[4:8].copy_(scal[5:9])
            dist.all_reduce(den, group=self.pg)
            gscale = float(self.world_size)
        for i, (fam, o) in enumerate((("rel", 1), ("act", 5))):
            sums = scal[o:o + 4] if gscale == 1.0 else self.scal_den[4 * i:4 * i + 4]
            if ce_sums is not None:
                sums = ce_sums[4 * i:4 * i + 4]
            res.append(ops.set_loss_bwd(out[fam + "_logits"], matches[i], self.eos, sums, gscale) if want_grad else None)
        return scal, res[0], res[1], res[2], matches

    @staticmethod
    def losses_from_scalars(s):
        """Host-side view of the device sums (one small D2H read)."""
        s = s.detach().float().cpu()
        if s.numel() > 10 and float(s[10]) != 0.0:
            # scipy raises ValueError("matrix contains invalid numeric entries") / the reference's prob[:, tgt_ids]
            # raises IndexError for these inputs (lxrt/matcher.py:73-79); the step's gradients are not to be trusted
            raise ValueError("matching refused %d frame(s)/row(s) of this step: NaN or -inf logits, more than 32 "
                             "targets in a frame, or a label outside [0, n_classes)" % int(s[10]))
        rel = float(s[1] / s[2]) if float(s[2]) > 0 else 0.0
        act = float(s[5] / s[6]) if float(s[6]) > 0 else 0.0
        rel_err = 100.0 - 100.0 * float(s[4] / s[3]) if float(s[3]) > 0 else 100.0
        act_err = 100.0 - 100.0 * float(s[8] / s[7]) if float(s[7]) > 0 else 100.0
        return dict(qa=float(s[0]), rel_ce=rel, act_ce=act, total=float(s[0]) + rel + act, rel_err=rel_err,
                    act_err=act_err)

    # ------------------------------------------------------------------ optimiser
    def _ensure_opt_state(self):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.store.master)
            self.exp_avg_sq = torch.zeros_like(self.store.master)

    def current_lr(self):
        if self.t_total != -1:
            return self.lr * warmup_linear(self.step_count / self.t_total, self.warmup)
        return self.lr

    def _push_lr(self):
        """lr of this step → device scalar (async copy from pinned memory; outside any captured graph)."""
        i = self._lr_slot
        if self._lr_events[i] is not None:
            self._lr_events[i].synchronize()              # the copy that last read this slot has executed
        self.lr_host[i] = self.current_lr()
        self.lr_dev.copy_(self.lr_host[i:i + 1], non_blocking=True)
        ev = self._lr_events[i] or torch.cuda.Event()
        ev.record()
        self._lr_events[i] = ev
        self._lr_slot = (i + 1) % self._lr_ring

    def optimizer_step(self, lr_on_device=False):
        st = self.store
        self._ensure_opt_state()
        lr = self.current_lr()
        sumsq = self.scal[9:10]
        n_live = st.n_live
        ops.call("shg_grad_sumsq", ops.ptr(st.grad), ops.ptr(st.seg_start), ops.ptr(st.seg_prefix), st.n_seg, n_live,
                 ops.ptr(sumsq), ops.ptr(self.sumsq_ws), ops.stream(), work=("hbm", 4.0 * n_live))
        ops.call("shg_bert_adam_step", ops.ptr(st.master), ops.ptr(st.grad), ops.ptr(self.exp_avg),
                 ops.ptr(self.exp_avg_sq), ops.ptr(st.shadow), st.n_matrix, ops.ptr(st.seg_start),
                 ops.ptr(st.seg_prefix), st.n_seg, n_live, ops.ptr(sumsq), self.max_norm, lr,
                 ops.ptr(self.lr_dev) if lr_on_device else None, self.b1, self.b2,
                 self.eps, self.wd, 1.0 / self.world_size, 1, ops.stream(), work=("hbm", 34.0 * n_live))
        self.engine.refresh_derived()
        if not lr_on_device:
            self.step_count += 1

    def all_reduce_grads(self):
        """The step's only collective: sum the flat live-gradient range over ranks (NCCL over NVLink); the
        1/world_size average is folded into the optimiser's grad_scale."""
        all_reduce_live_grads(self.store, self.world_size, self.pg)

    # ------------------------------------------------------------------ steps
    def _advance_seed(self):
        if self.engine.dropout_on:
            self.engine.drop_seed.add_(1)

    def _reduce_async(self, slices):
        """Start the all-reduce of gradient-buffer slices whose contents are final once the kernels queued so far
        (main stream and weight-gradient stream) have run.  The collective is issued from the weight-gradient
        stream: NCCL's stream then waits for that stream's queue (ordered after the main stream's current point by
        one event) while the main stream goes on with the backward — nothing drains."""
        import torch.distributed as dist
        eng = self.engine
        cur = torch.cuda.current_stream()
        if not eng.use_streams:
            return [dist.all_reduce(self.store.grad[sl], group=self.pg, async_op=True) for sl in slices]
        eng.sw.wait_event(cur.record_event())
        with torch.cuda.stream(eng.sw):
            return [dist.all_reduce(self.store.grad[sl], group=self.pg, async_op=True) for sl in slices]

    def train_step(self, inp, batch, lr_on_device=False):
        """One training step.  Data-parallel runs (world_size > 1) overlap the path's only collective with the
        backward: the gradient span is reduced in groups (ParamStore.dp_groups), each started as soon as its
        gradients are final — everything after the LXRT encoder (decoders, heads, CrossEncoder, hoisted K/V weights)
        under the whole LXRT backward; each relation layer right after its own backward (the language layers and
        embedding tables ride with relation layer 2); only relation layer 0 and the small vectors (30 MB) are left
        for the end.  The same code path runs eagerly and under CUDA-graph capture (the NCCL kernels become nodes of
        the step's graph on their own branch)."""
        self._advance_seed()
        out, ctx = self.engine.forward(inp, train=True)
        scal, d_hg, d_rel, d_act, _ = self.loss_and_grads(out, batch)
        if self.task != "hgqa":
            self.engine.backward(ctx, None, None, None, need_dvisn=self.backbone_grad, d_logit=d_hg)
            if self.world_size > 1:
                self.all_reduce_grads()
        elif self.world_size == 1:
            self.engine.backward(ctx, d_hg, d_rel, d_act, need_dvisn=self.backbone_grad)
        else:
            # the path's only collective, overlapped: every group of gradients is reduced as soon as it is final
            eng, grp = self.engine, self.store.dp_groups
            eng.backward_late_params(ctx, d_hg, d_rel, d_act)
            first = list(grp["late"])
            lang_early = bool(ctx.get("_lang_started"))
            if lang_early:
                # the language side was started under the decoders' backward (Engine.backward_lang_start) and has
                # finished by now: its matrices and the embedding tables go out with the first group
                eng.backward_lang_join(ctx)
                first += grp["lang"] + grp["emb"]
            works = self._reduce_async(first)
            # from here to the end of the backward an all-reduce is in flight most of the time and owns some SMs:
            # the persistent GEMM grids leave those to it (include/shgvqa.h: shg_gemm_args.max_ctas)
            with ops.gemm_cta_cap(self.dp_gemm_ctas):
                eng.backward_early_begin(ctx)
                for i in reversed(range(eng.nr)):
                    eng.backward_r_layer(ctx, i, self.backbone_grad)
                    if i == 0:
                        break
                    slices = list(grp["r%d" % i])
                    if i == min(2, eng.nr - 1) and not lang_early:
                        eng.backward_lang_join(ctx)            # the 40-token language side finished long ago
                        slices += grp["lang"] + grp["emb"]
                    works += self._reduce_async(slices)
                eng.backward_early_end(ctx, self.backbone_grad)
            works += self._reduce_async(grp["r0"] + grp.get("front", []) + grp["vec"])
            for w in works:
                w.wait()                                   # stream-level: the optimiser waits for NCCL's stream
        self.d_feat = ctx.get("d_feat")      # gradient wrt the backbone features (backbone_grad runs)
        self.optimizer_step(lr_on_device)
        return out, scal

    # ------------------------------------------------------------------ CUDA-graph step
    def _snapshot_state(self):
        st, live = self.store, self.store.live_extent
        return dict(master=st.master[live].clone(), m=self.exp_avg[live].clone(), v=self.exp_avg_sq[live].clone(),
                    seed=self.engine.drop_seed.clone(), step=self.step_count)

    def _restore_state(self, snap):
        st, live = self.store, self.store.live_extent
        st.master[live].copy_(snap["master"])
        self.exp_avg[live].copy_(snap["m"])
        self.exp_avg_sq[live].copy_(snap["v"])
        self.engine.drop_seed.copy_(snap["seed"])
        self.step_count = snap["step"]
        st.grad.zero_()
        st.refresh_shadow()
        self.engine.refresh_derived()

    def capture(self, inp, batch, warmup=3):
        """Capture one whole training step (≈650 launches on several streams; for world_size > 1 the three NCCL
        all-reduces too) into ONE CUDA graph over STATIC input buffers.  The eager path is launch-bound on the
        host; the replay is one launch.  lr is read from device memory so the warm-up schedule keeps working; new
        batches are copied into the static buffers before each replay.
        Capturing has no side effect on the training state: the eager steps that warm the allocator, the
        TMA-descriptor cache and NCCL are real optimiser steps, so weights, Adam moments, the dropout seed and the
        step count are snapshotted before and restored after (the reference's trajectory, tasks/agqaHGQA.py:390-392,
        continues exactly where it was)."""
        self._ensure_opt_state()
        from . import _lib
        snap = self._snapshot_state()
        l0 = _lib.launch_count
        self.train_step(inp, batch)
        self.launches_per_step = _lib.launch_count - l0
        self.static_inp = {k: v.clone() for k, v in inp.items()}
        self.static_batch = {k: v.clone() for k, v in batch.items()}
        self._push_lr()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                       # warm allocator, TMA-descriptor cache, cuFuncSetAttribute
                self.train_step(self.static_inp, self.static_batch, lr_on_device=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread keeps polling its events while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.static_out, self.static_scal = self.train_step(self.static_inp, self.static_batch,
                                                                lr_on_device=True)
        self._restore_state(snap)
        torch.cuda.synchronize()
        return self.graph

    def release(self):
        """Drop the captured graph (it holds NCCL resources: destroy_process_group() blocks while it is alive)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def train_step_graph(self, inp=None, batch=None):
        """Replay the captured step; optional new inputs (device or pinned-host tensors) are copied in first.
        CSR label buffers must keep the captured sizes (pad the label vector, offsets carry the real extents)."""
        if inp is not None:
            for k, v in inp.items():
                self.static_inp[k].copy_(v, non_blocking=True)
        if batch is not None:
            for k, v in batch.items():
                self.static_batch[k].copy_(v, non_blocking=True)
        self._push_lr()
        self.graph.replay()
        self.step_count += 1
        return self.static_out, self.static_scal

    def prefetch_host(self, host_inp, host_batch):
        """Start the H2D copies of the NEXT step's inputs (pinned host tensors) on the copy stream into staging
        buffers; they overlap whatever the compute streams are running.  The next train_step_host consumes them
        with a device-to-device copy instead of waiting for PCIe."""
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(self.device)
            self.stage_inp = {k: torch.empty_like(v) for k, v in self.static_inp.items()}
            self.stage_batch = {k: torch.empty_like(v) for k, v in self.static_batch.items()}
            self.stage_ready = torch.cuda.Event()
            self.stage_free = torch.cuda.Event()
            self.stage_free.record()
        self.copy_stream.wait_event(self.stage_free)          # the previous consumer has read the staging buffers
        with torch.cuda.stream(self.copy_stream):
            for k, v in host_inp.items():
                self.stage_inp[k].copy_(v, non_blocking=True)
            for k, v in host_batch.items():
                self.stage_batch[k].copy_(v, non_blocking=True)
            self.stage_ready.record()
        self._staged = True

    def drain_losses(self):
        """Loss dict of the most recent lagged step whose result has not been handed out yet (waits for its D2H
        copy), or None."""
        if self._loss_slots is None:
            return None
        j = self._loss_idx ^ 1
        if not self._loss_pending[j]:
            return None
        self._loss_events[j].synchronize()
        self._loss_pending[j] = False
        return self.losses_from_scalars(self._loss_slots[j])

    def train_step_host(self, host_inp, host_batch, next_inp=None, next_batch=None, lagged=False):
        """The call a user of the path makes with HOST (pinned) buffers: H2D copies of the step's inputs, the
        step, and a D2H read of the loss sums.  Returns the loss dict.  next_inp / next_batch (the following
        step's host buffers, as a data loader has them) are prefetched to the device while this step computes."""
        if self.graph is not None:
            if self._staged:
                cur = torch.cuda.current_stream()
                cur.wait_event(self.stage_ready)
                for k in host_inp:
                    self.static_inp[k].copy_(self.stage_inp[k], non_blocking=True)
                for k in host_batch:
                    self.static_batch[k].copy_(self.stage_batch[k], non_blocking=True)
                self.stage_free.record()
                self._staged = False
                _, scal = self.train_step_graph()
            else:
                _, scal = self.train_step_graph(host_inp, host_batch)
            if next_inp is not None:
                self.prefetch_host(next_inp, next_batch)
            if not lagged:
                return self.losses_from_scalars(scal)
            # lagged read: this step's loss sums start their D2H copy into a pinned slot now and are handed out by the
            # NEXT call (or drain_losses), so the host enqueues step i+1 while step i still runs — every step's
            # result is still read back, just not synchronously
            if self._loss_slots is None:
                self._loss_slots = [torch.zeros(16, dtype=torch.float32).pin_memory() for _ in range(2)]
                self._loss_events = [torch.cuda.Event(), torch.cuda.Event()]
                self._loss_pending = [False, False]
                self._loss_idx = 0
            i = self._loss_idx
            prev = self.drain_losses()
            self._loss_slots[i].copy_(scal, non_blocking=True)
            self._loss_events[i].record()
            self._loss_pending[i] = True
            self._loss_idx = i ^ 1
            return prev
        inp = {k: v.to(self.device, non_blocking=True) for k, v in host_inp.items()}
        batch = {k: v.to(self.device, non_blocking=True) for k, v in host_batch.items()}
        _, scal = self.train_step(inp, batch)
        return self.losses_from_scalars(scal)

    @torch.no_grad()
    def eval_step(self, inp):
        out, _ = self.engine.forward(inp, train=False)
        return out
