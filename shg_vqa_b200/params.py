"""Parameter store with the reference's exact ``state_dict`` layout.

The reference saves ``model.state_dict()`` and reloads it with ``strict=True``
(tasks/agqaHGQA.py:859-874).  That key set — 669 entries for --taskHGQA, including
weight ALIASES (one CrossLayer registered as ``cross_attn_layer.cross``,
``x_layers.0`` and ``x_layers.1``; ``pooler`` ≡ ``pooler_dict.cross``;
lxrt/modeling_capsbert.py:1122-1124, 1247-1249, 1144, 1811) and dead-but-serialised
branches (``self``, ``cross_self``, ``old`` layers, unused poolers, ``box_fc``) — is a
hard boundary, so it is reproduced here by construction and checked against
tests/golden/state_dict_keys.json.

Storage is B200-first rather than module-first: ONE flat fp32 master buffer, ONE
flat fp32 gradient buffer (what the NCCL all-reduce and the fused optimiser walk)
and ONE flat bf16 shadow of the matrices the tensor cores read.  The flat order
puts query/key/value weights (and biases) of each attention block next to each
other so the fused [2304,768] QKV projection reads them in place.
"""
import collections

import torch

H, I_FFN, D_FFN = 768, 3072, 2048


def _bert_attention(p):
    out = []
    for n in ("query", "key", "value"):
        out += [(p + "." + n + ".weight", (H, H)), (p + "." + n + ".bias", (H,))]
    return out


def _att_output(p):
    return [(p + ".dense.weight", (H, H)), (p + ".dense.bias", (H,)),
            (p + ".LayerNorm.weight", (H,)), (p + ".LayerNorm.bias", (H,))]


def _ffn(p_inter, p_out):
    return [(p_inter + ".dense.weight", (I_FFN, H)), (p_inter + ".dense.bias", (I_FFN,)),
            (p_out + ".dense.weight", (H, I_FFN)), (p_out + ".dense.bias", (H,)),
            (p_out + ".LayerNorm.weight", (H,)), (p_out + ".LayerNorm.bias", (H,))]


def _cross_layer(p):            # CrossLayer, modeling_capsbert.py:624-656
    return (_bert_attention(p + ".visual_attention.att") + _att_output(p + ".visual_attention.output")
            + _ffn(p + ".lang_inter", p + ".lang_output") + _ffn(p + ".visn_inter", p + ".visn_output"))


def _self_cross_layer(p):       # SelfCrossLayer (dead branch), :679-
    return (_bert_attention(p + ".cross_att.self") + _att_output(p + ".cross_att.output")
            + _ffn(p + ".vl_inter", p + ".vl_output"))


def _cross_and_self_layer(p):   # CrossAndSelfLayer (dead branch)
    return (_bert_attention(p + ".visual_attention.att") + _att_output(p + ".visual_attention.output")
            + _bert_attention(p + ".self_att_layer.self") + _att_output(p + ".self_att_layer.output")
            + _ffn(p + ".vl_inter", p + ".vl_output"))


def _bert_layer(p):             # BertLayer, :492-503
    return (_bert_attention(p + ".attention.self") + _att_output(p + ".attention.output")
            + _ffn(p + ".intermediate", p + ".output"))


def _attn_dict(p):
    return (_cross_layer(p + ".cross") + _self_cross_layer(p + ".self")
            + _cross_and_self_layer(p + ".cross_self") + _cross_layer(p + ".old"))


def _poolers(p):
    return [(p + "_dict.cross.dense2.weight", (H, 2 * H)), (p + "_dict.cross.dense2.bias", (H,)),
            (p + "_dict.self.dense.weight", (H, H)), (p + "_dict.self.dense.bias", (H,)),
            (p + "_dict.cross_self.dense.weight", (H, H)), (p + "_dict.cross_self.dense.bias", (H,)),
            (p + "_dict.no_cross.dense2.weight", (H, 2 * H)), (p + "_dict.no_cross.dense2.bias", (H,)),
            (p + "_dict.old.dense.weight", (H, H)), (p + "_dict.old.dense.bias", (H,)),
            (p + ".dense2.weight", (H, 2 * H)), (p + ".dense2.bias", (H,))]


def _decoder_layer(p):          # TransformerDecoderLayer, transformer.py:187-210
    out = []
    for a in ("self_attn", "multihead_attn"):
        out += [(p + "." + a + ".in_proj_weight", (3 * H, H)), (p + "." + a + ".in_proj_bias", (3 * H,)),
                (p + "." + a + ".out_proj.weight", (H, H)), (p + "." + a + ".out_proj.bias", (H,))]
    out += [(p + ".linear1.weight", (D_FFN, H)), (p + ".linear1.bias", (D_FFN,)),
            (p + ".linear2.weight", (H, D_FFN)), (p + ".linear2.bias", (H,))]
    for n in ("norm1", "norm2", "norm3"):
        out += [(p + "." + n + ".weight", (H,)), (p + "." + n + ".bias", (H,))]
    return out


def _mlp_head(p, n_out):        # agqa_model.py:105-110
    return [(p + ".0.weight", (2 * H, H)), (p + ".0.bias", (2 * H,)), (p + ".2.weight", (2 * H,)),
            (p + ".2.bias", (2 * H,)), (p + ".3.weight", (n_out, 2 * H)), (p + ".3.bias", (n_out,))]


def _hg_embed(p, n_queries):
    return [(p + ".word_embeddings.weight", (n_queries, H)), (p + ".token_type_embeddings.weight", (16, H)),
            (p + ".LayerNorm.weight", (H,)), (p + ".LayerNorm.bias", (H,))]


def build_key_shapes(task="hgqa", n_answers=171, n_rel_cls=457, n_act_cls=158, n_rel_q=128, n_act_q=48,
                     l_layers=5, x_layers=2, r_layers=5, d_layers=5, vocab=30522):
    """Ordered {key: (shape, dtype)} equal to the reference AGQAModel's state_dict (backbone excluded)."""
    e = "lxrt_encoder.model.bert"
    ks = [(e + ".embeddings.word_embeddings.weight", (vocab, H)),
          (e + ".embeddings.position_embeddings.weight", (512, H)),
          (e + ".embeddings.token_type_embeddings.weight", (2, H)),
          (e + ".embeddings.LayerNorm.weight", (H,)), (e + ".embeddings.LayerNorm.bias", (H,)),
          (e + ".encoder.visn_fc.cls_token", (1, 1, H)),
          (e + ".encoder.visn_fc.conv.1.weight", (H, 2048, 5, 3, 3)), (e + ".encoder.visn_fc.conv.1.bias", (H,)),
          (e + ".encoder.visn_fc.conv.4.weight", (H, H, 5, 3, 3)), (e + ".encoder.visn_fc.conv.4.bias", (H,)),
          (e + ".encoder.visn_fc.position_encoding.position_ids", (1, 393)),
          (e + ".encoder.visn_fc.position_encoding.pe.weight", (393, H)),
          (e + ".encoder.visn_fc.box_fc.weight", (H, 4)), (e + ".encoder.visn_fc.box_fc.bias", (H,)),
          (e + ".encoder.visn_fc.pos_layer_norm.weight", (H,)), (e + ".encoder.visn_fc.pos_layer_norm.bias", (H,))]
    ks += _attn_dict(e + ".encoder.cross_attn_layer")
    for i in range(l_layers):
        ks += _bert_layer("%s.encoder.layer.%d" % (e, i))
    for i in range(x_layers):
        ks += _cross_layer("%s.encoder.x_layers.%d" % (e, i))
    for i in range(r_layers):
        ks += _bert_layer("%s.encoder.r_layers.%d" % (e, i))
    ks += _poolers(e + ".pooler")
    if task == "hgqa":
        g = "hgq_encoder"
        ks += [(g + ".act_token", (1, 1, H)), (g + ".rel_token", (1, 1, H)), (g + ".cls_token", (1, 1, H))]
        ks += _attn_dict(g + ".cross_attn_layer")
        for i in range(x_layers):
            ks += _cross_layer("%s.x_layers.%d" % (g, i))
        ks += _poolers(g + ".pooler")
        ks += _hg_embed("relation_query_embed", n_rel_q) + _hg_embed("action_query_embed", n_act_q)
        for i in range(d_layers):
            ks += _decoder_layer("rel_decoder.layers.%d" % i)
        ks += _mlp_head("class_embed", n_rel_cls)
        for i in range(d_layers):
            ks += _decoder_layer("action_decoder.layers.%d" % i)
        ks += _mlp_head("action_embed", n_act_cls)
    ks += _mlp_head("logit_fc", n_answers)
    out = collections.OrderedDict()
    for k, shape in ks:
        out[k] = (tuple(shape), "int64" if k.endswith("position_ids") else "float32")
    return out


def canonical_key(key):
    """Alias → the one name that owns the storage."""
    for n in range(8):
        key = key.replace(".x_layers.%d." % n, ".cross_attn_layer.cross.")
    if ".pooler.dense2." in key:
        key = key.replace(".pooler.dense2.", ".pooler_dict.cross.dense2.")
    return key


def _is_conv_weight(key, shape):
    """Conv3d weights of the visual front-end: kept PACKED [c_out, 5, 3, 3, c_in] in the flat buffers (the layout
    the implicit-GEMM kernels read, csrc/conv.cu); p()/g()/state_dict() expose the reference's
    [c_out, c_in, 5, 3, 3] as a permuted view of that storage."""
    return len(shape) == 5 and ".visn_fc.conv." in key


def _is_matrix(key, shape):
    """Tensors that get a bf16 shadow (tensor-core operands)."""
    if _is_conv_weight(key, shape):
        return True
    return len(shape) == 2 and shape[1] >= 64 and "embeddings" not in key and "pe.weight" not in key \
        and "box_fc" not in key


class ParamStore:
    """Flat fp32 master / fp32 grad / bf16 shadow buffers with named views."""

    def __init__(self, key_shapes, device, with_grad=True, live=None):
        """live: key -> bool, "receives a gradient in training" (None: everything).  The flat order is
        [frozen matrices | live matrices | live vectors+tables | frozen rest], so the live range — what the
        all-reduce and the optimiser walk — is ONE contiguous span and the bf16 shadow is the leading span."""
        self.key_shapes = key_shapes
        self.device = torch.device(device)
        canon = collections.OrderedDict()
        for k, (shape, dtype) in key_shapes.items():
            if dtype == "float32":
                canon.setdefault(canonical_key(k), tuple(shape))
        is_live = (lambda k: True) if live is None else live
        mats = [k for k, s in canon.items() if _is_matrix(k, s)]
        rest = [k for k, s in canon.items() if not _is_matrix(k, s)]
        # q|k|v weights (and biases) stay adjacent inside each group: key order is preserved within a group
        order = ([k for k in mats if not is_live(k)] + [k for k in mats if is_live(k)]
                 + [k for k in rest if is_live(k)] + [k for k in rest if not is_live(k)])
        self.offsets = {}
        self.live_keys = [k for k in order if is_live(k)]
        off = 0
        self.live_begin = self.live_end = None
        n_mats = len(mats)
        for idx, k in enumerate(order):
            if idx == n_mats:
                self.n_matrix = off
            if is_live(k) and self.live_begin is None:
                self.live_begin = off
            n = 1
            for d in canon[k]:
                n *= d
            self.offsets[k] = (off, n, canon[k])
            off += (n + 7) // 8 * 8          # 32-byte aligned fp32 / 16-byte aligned bf16 views
            if is_live(k):
                self.live_end = off
        if n_mats == len(order):
            self.n_matrix = off
        self.n_total = off
        self.n_live = self.live_end - self.live_begin
        self.live_extent = slice(self.live_begin, self.live_end)
        self.master = torch.zeros(self.n_total, device=self.device, dtype=torch.float32)
        self.grad = torch.zeros(self.n_total, device=self.device, dtype=torch.float32) if with_grad else None
        self.shadow = torch.zeros(self.n_matrix, device=self.device, dtype=torch.bfloat16) \
            if self.device.type == "cuda" else None
        self.int_buffers = {k: torch.arange(s[-1], device=self.device).expand(s).clone()
                            for k, (s, d) in key_shapes.items() if d == "int64"}
        # data-parallel buckets over the live span.  "late" = tensors whose gradients are final after the first half
        # of the backward (everything that is not the LXRT encoder's language/relation layers and embeddings):
        # the live matrices from the first such key to the end of the matrix region.  The rest of the live span
        # ("early" parameters: language/relation layers, and every vector/table) is reduced at the end.
        late_mats = [k for k in mats if is_live(k) and not k.startswith("lxrt_encoder.")]
        self.bucket_late = slice(self.offsets[late_mats[0]][0], self.n_matrix) if late_mats else slice(0, 0)
        self.bucket_early_a = slice(self.live_begin, self.bucket_late.start if late_mats else self.live_begin)
        self.bucket_early_b = slice(self.n_matrix if late_mats else self.live_begin, self.live_end)
        # finer split of the early parameters for the data-parallel schedule: the matrices of the UPPER language /
        # relation layers (index >= split) are final after the first part of the second backward half and are
        # reduced under the rest of it
        self.bucket_early_upper, self.bucket_early_rest = self._split_early(mats, is_live, upper_from=2)
        self.dp_groups = self._dp_groups(is_live)
        # optimiser segment table (one contiguous live span)
        self.n_seg = 1
        self.seg_start = torch.tensor([self.live_begin], dtype=torch.int64, device=self.device)
        self.seg_prefix = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._shadow_fresh = False

    def _dp_groups(self, is_live):
        """Live span -> {group: [slices]} for the overlapped data-parallel schedule (HotPathTrainer.train_step):
        'late' (matrices of everything after the LXRT encoder: final after the first backward half), 'r0'..'r4' (one
        relation layer's matrices each: final right after that layer's backward), 'lang' (language-layer matrices)
        and 'emb' (BertEmbeddings tables): final once the language side stream has run, 'front' (Conv3d front-end),
        'vec' (every bias / LayerNorm vector / token: final at the very end).  The groups tile the live span."""
        import re
        pat = re.compile(r"\.encoder\.(layer|r_layers)\.(\d+)\.")
        groups = collections.OrderedDict()
        order = sorted(self.live_keys, key=lambda k: self.offsets[k][0])
        for i, k in enumerate(order):
            off, n, shape = self.offsets[k]
            end = self.offsets[order[i + 1]][0] if i + 1 < len(order) else self.live_end
            m = pat.search(k)
            if ".visn_fc." in k:
                g = "front"
            elif off >= self.n_matrix:
                g = "emb" if k.startswith("lxrt_encoder.model.bert.embeddings.") and k.endswith("embeddings.weight") else "vec"
            elif not k.startswith("lxrt_encoder."):
                g = "late"
            elif m and m.group(1) == "layer":
                g = "lang"
            elif m:
                g = "r%d" % int(m.group(2))
            else:
                g = "late"           # LXRT cross layers / pooler when they are live (--taskVQA, x_grad)
            lst = groups.setdefault(g, [])
            if lst and lst[-1].stop == off:
                lst[-1] = slice(lst[-1].start, end)
            else:
                lst.append(slice(off, end))
        return groups

    def _split_early(self, mats, is_live, upper_from):
        """(upper, rest) slice lists over the early parameters.  upper = what is final after part 0 of the second
        backward half (engine.backward_early_params): every language-layer matrix, the relation-layer matrices with
        index >= upper_from, and the BertEmbeddings tables (the language side is 40 tokens per clip and finishes long
        before the relation side); rest = the lower relation layers, the Conv3d front-end and the small vectors."""
        import re
        pat = re.compile(r"\.encoder\.(layer|r_layers)\.(\d+)\.")
        upper, rest = [], []
        lo, hi = self.bucket_early_a.start, self.bucket_early_a.stop

        def push(lst, a, b):
            if b > a:
                if lst and lst[-1].stop == a:
                    lst[-1] = slice(lst[-1].start, b)
                else:
                    lst.append(slice(a, b))
        keys = [k for k in mats if is_live(k) and lo <= self.offsets[k][0] < hi]
        if keys:
            push(rest, lo, self.offsets[keys[0]][0])
        for i, k in enumerate(keys):
            a = self.offsets[k][0]
            b = self.offsets[keys[i + 1]][0] if i + 1 < len(keys) else hi
            m = pat.search(k)
            is_upper = bool(m) and (m.group(1) == "layer" or int(m.group(2)) >= upper_from)
            push(upper if is_upper else rest, a, b)
        # vectors / tables: the BertEmbeddings block is the leading run of the region (key order)
        a = b = self.bucket_early_b.start
        for k in self.live_keys:
            off, n, _ = self.offsets[k]
            if off == b and k.startswith("lxrt_encoder.model.bert.embeddings."):
                b = off + (n + 7) // 8 * 8
        push(upper, a, b)
        push(rest, b, self.bucket_early_b.stop)
        return upper, rest

    # ------------------------------------------------------------------ views
    def _view(self, buf, key):
        off, n, shape = self.offsets[canonical_key(key)]
        if _is_conv_weight(key, shape):
            return buf[off:off + n].view(shape[0], shape[2], shape[3], shape[4], shape[1]).permute(0, 4, 1, 2, 3)
        return buf[off:off + n].view(shape)

    def packed(self, buf, key):
        """Conv3d weight (or its gradient / shadow) as the packed matrix [c_out, 45 * c_in] it is stored as."""
        off, n, shape = self.offsets[canonical_key(key)]
        assert _is_conv_weight(key, shape), key
        return buf[off:off + n].view(shape[0], n // shape[0])

    def p(self, key):
        return self._view(self.master, key)

    def g(self, key):
        return self._view(self.grad, key)

    def w(self, key):
        """bf16 shadow of a matrix."""
        off, n, shape = self.offsets[canonical_key(key)]
        assert off + n <= self.n_matrix, key + " has no bf16 shadow"
        if _is_conv_weight(key, shape):
            return self.packed(self.shadow, key)
        return self.shadow[off:off + n].view(shape)

    def fused(self, buf, first_key, rows, cols=None):
        """View spanning several adjacent tensors, e.g. query|key|value as one [2304,768] matrix."""
        off, _, shape = self.offsets[canonical_key(first_key)]
        if cols is None:
            return buf[off:off + rows]
        return buf[off:off + rows * cols].view(rows, cols)

    def refresh_shadow(self):
        from . import ops
        ops.to_bf16(self.master[:self.n_matrix], out=self.shadow)
        self._shadow_fresh = True

    # ------------------------------------------------------------------ reference-compatible I/O
    def state_dict(self):
        out = collections.OrderedDict()
        for k, (shape, dtype) in self.key_shapes.items():
            out[k] = self.int_buffers[k] if dtype == "int64" else self.p(k)
        return out

    def load_state_dict(self, sd, strict=True):
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}   # agqaHGQA.py:866-869
        missing = [k for k in self.key_shapes if k not in sd]
        unexpected = [k for k in sd if k not in self.key_shapes and not k.startswith("vid_encoder.")]
        if strict and (missing or unexpected):
            raise RuntimeError("state_dict mismatch: missing %s unexpected %s" % (missing[:5], unexpected[:5]))
        for k, (shape, dtype) in self.key_shapes.items():
            if k not in sd:
                continue
            v = sd[k]
            if tuple(v.shape) != tuple(shape):
                raise RuntimeError("size mismatch for %s: %s vs %s" % (k, tuple(v.shape), tuple(shape)))
            if dtype == "int64":
                self.int_buffers[k].copy_(v)
            else:
                self.p(k).copy_(v.to(self.device, torch.float32))
        self._shadow_fresh = False
        if self.shadow is not None:
            self.refresh_shadow()
        return missing, unexpected
