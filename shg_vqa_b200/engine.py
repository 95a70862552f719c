"""Forward and hand-written backward of the SHG-VQA hot path over libshgvqa.

This is the host-side schedule of the path SURVEY.md §8a names — embeddings (E0),
language / relation BertLayers (L, R), the weight-shared CrossLayers (X), the two
DETR decoders over per-frame relation / action queries (D), the MLP heads (H),
the hypergraph x question CrossEncoder (HG), the poolers (P), the per-frame
Hungarian matching and the set loss (M1-M4) — expressed as a flat sequence of
C-ABI launches on the current CUDA stream.  No torch autograd, no torch math on
the path: torch owns buffers and streams only, so a whole step is capturable in
one CUDA graph.

B200-first choices (vs the reference's module-by-module eager fp32):
  * q|k|v (and the decoders' packed in_proj) are one [2304,768] projection;
  * the ten decoder layers' memory K/V projections are hoisted into ONE
    [B*393,768] x [768,15360] GEMM (and one dgrad on the way back);
  * bias, GELU/ReLU/tanh, residual adds and the backward-through-activation are GEMM
    epilogues; LayerNorm backward also emits the bias gradient of the Linear before it;
  * forward/dgrad/wgrad all read operands where they lie (MN-major UMMA descriptors),
    wgrad accumulates with split-K fp32 atomics straight into the flat gradient buffer;
  * matching + weighted CE + its gradient never leave the device.
"""
import os

import torch

from . import ops
from .params import ParamStore

BF16 = torch.bfloat16
H = 768
EPS_BERT = 1e-12     # lxrt/modeling_capsbert.py:310,338,428,482 ; tasks/agqa_model.py:108
EPS_DETR = 1e-5      # lxrt/transformer.py:199-201


class Lin:
    """One projection: bf16 shadow weight [N,K], fp32 bias [N], fp32 gradient views."""

    def __init__(self, w, b, gw, gb):
        self.w, self.b, self.gw, self.gb = w, b, gw, gb


class Norm:
    def __init__(self, g, b, gg, gb):
        self.g, self.b, self.gg, self.gb = g, b, gg, gb


class Engine:
    def __init__(self, store: ParamStore, task="hgqa", n_frames=16, n_rel=8, n_act=3, l_layers=5, r_layers=5,
                 x_layers=2, d_layers=5):
        self.st = store
        self.task = task
        self.T, self.n_rel, self.n_act = n_frames, n_rel, n_act
        self.nl, self.nr, self.nx, self.nd = l_layers, r_layers, x_layers, d_layers
        self.dev = store.device
        self._build_handles()
        # training-mode dropout (reference probabilities: hidden / attention-prob 0.1, lxrt/modeling_capsbert.py:216-217;
        # decoders and action-query embedding 0.15, param.py:152-153; relation-query embedding 0.1, :302).  Off by
        # default: parity is stated for eval()/p=0 semantics; the trainer switches it on for training runs.
        self.dropout_on = False
        self.p_hidden, self.p_attn, self.p_dec, self.p_emb, self.p_relq, self.p_actq = 0.1, 0.1, 0.15, 0.1, 0.1, 0.15
        self.drop_seed = torch.zeros(1, device=self.dev, dtype=torch.int32) if self.dev.type == "cuda" else None
        self._salt, self._training = 0, False
        self.nv_mem = 393
        # Independent sub-graphs run on side streams so that the small-M launches (40-token language side,
        # 48-query action decoder) fill SMs the 148-CTA persistent kernels of the visual side leave idle.
        self.use_streams = self.dev.type == "cuda"
        import os
        self.fused_decoders = os.environ.get("SHG_FUSED_DECODERS", "1") != "0"
        # (stream priorities — chains high, weight gradients low — were measured: no effect on the replayed step)
        self.s1 = torch.cuda.Stream(self.dev) if self.use_streams else None
        self.s2 = torch.cuda.Stream(self.dev) if self.use_streams else None
        # weight/bias-gradient products are off the critical path (nothing reads them before the optimiser): they
        # are queued on their own stream and fill the SMs the dependent chain of small launches leaves idle
        self.sw = torch.cuda.Stream(self.dev) if self.use_streams else None
        self._keep = []          # operands of in-flight wgrads stay referenced until the final join
        self.early_lang_bwd = os.environ.get("SHG_EARLY_LANG_BWD", "1") != "0"   # see backward_lang_start

    def _fork(self, *streams):
        """streams start after everything queued so far on the current stream."""
        if not self.use_streams:
            return
        ev = torch.cuda.current_stream().record_event()
        for st in streams:
            st.wait_event(ev)

    def _join(self, *streams):
        """the current stream waits for the given streams."""
        if not self.use_streams:
            return
        cur = torch.cuda.current_stream()
        for st in streams:
            cur.wait_event(st.record_event())

    def _on(self, st):
        """Run the enclosed launches on side stream `st` (None: stay on the current stream), in that stream's GEMM
        workspace lane (ops.workspace_lane)."""
        import contextlib
        if not (self.use_streams and st is not None):
            return contextlib.nullcontext()
        stack = contextlib.ExitStack()
        stack.enter_context(torch.cuda.stream(st))
        stack.enter_context(ops.workspace_lane(1 if st is self.s1 else (2 if st is self.s2 else 3)))
        return stack

    # ------------------------------------------------------------------ parameter handles
    def _lin(self, wkey, bkey=None):
        st = self.st
        bkey = bkey or wkey.replace(".weight", ".bias")
        return Lin(st.w(wkey), st.p(bkey), st.g(wkey) if st.grad is not None else None,
                   st.g(bkey) if st.grad is not None else None)

    def _norm(self, prefix):
        st = self.st
        g = st.grad is not None
        return Norm(st.p(prefix + ".weight"), st.p(prefix + ".bias"),
                    st.g(prefix + ".weight") if g else None, st.g(prefix + ".bias") if g else None)

    def _qkv(self, prefix):
        """query|key|value of a BertAttention as one [2304,768] projection (adjacent in the flat buffers)."""
        st = self.st
        g = st.grad is not None
        return Lin(st.fused(st.shadow, prefix + ".query.weight", 3 * H, H),
                   st.fused(st.master, prefix + ".query.bias", 3 * H),
                   st.fused(st.grad, prefix + ".query.weight", 3 * H, H) if g else None,
                   st.fused(st.grad, prefix + ".query.bias", 3 * H) if g else None)

    def _rows(self, lin, r0, r1):
        return Lin(lin.w[r0:r1], lin.b[r0:r1], lin.gw[r0:r1] if lin.gw is not None else None,
                   lin.gb[r0:r1] if lin.gb is not None else None)

    def _bert_layer(self, p):
        return dict(qkv=self._qkv(p + ".attention.self"), o=self._lin(p + ".attention.output.dense.weight"),
                    ln1=self._norm(p + ".attention.output.LayerNorm"),
                    f1=self._lin(p + ".intermediate.dense.weight"), f2=self._lin(p + ".output.dense.weight"),
                    ln2=self._norm(p + ".output.LayerNorm"))

    def _cross_layer(self, p):
        a = p + ".visual_attention"
        return dict(qkv=self._qkv(a + ".att"), o=self._lin(a + ".output.dense.weight"),
                    lno=self._norm(a + ".output.LayerNorm"),
                    lf1=self._lin(p + ".lang_inter.dense.weight"), lf2=self._lin(p + ".lang_output.dense.weight"),
                    lln=self._norm(p + ".lang_output.LayerNorm"),
                    vf1=self._lin(p + ".visn_inter.dense.weight"), vf2=self._lin(p + ".visn_output.dense.weight"),
                    vln=self._norm(p + ".visn_output.LayerNorm"))

    def _decoder_layer(self, p):
        sa = self._lin(p + ".self_attn.in_proj_weight", p + ".self_attn.in_proj_bias")
        ca = self._lin(p + ".multihead_attn.in_proj_weight", p + ".multihead_attn.in_proj_bias")
        return dict(sa_qk=self._rows(sa, 0, 2 * H), sa_v=self._rows(sa, 2 * H, 3 * H),
                    sa_o=self._lin(p + ".self_attn.out_proj.weight"),
                    ca_q=self._rows(ca, 0, H), ca_kv=self._rows(ca, H, 3 * H),
                    ca_o=self._lin(p + ".multihead_attn.out_proj.weight"),
                    l1=self._lin(p + ".linear1.weight"), l2=self._lin(p + ".linear2.weight"),
                    n1=self._norm(p + ".norm1"), n2=self._norm(p + ".norm2"), n3=self._norm(p + ".norm3"))

    def _head(self, p):
        return dict(l0=self._lin(p + ".0.weight"), ln=self._norm(p + ".2"), l3=self._lin(p + ".3.weight"))

    def _build_handles(self):
        st = self.st
        e = "lxrt_encoder.model.bert"
        g = st.grad is not None
        self.emb = dict(word=st.p(e + ".embeddings.word_embeddings.weight"),
                        pos=st.p(e + ".embeddings.position_embeddings.weight"),
                        type=st.p(e + ".embeddings.token_type_embeddings.weight"),
                        ln=self._norm(e + ".embeddings.LayerNorm"),
                        gword=st.g(e + ".embeddings.word_embeddings.weight") if g else None,
                        gpos=st.g(e + ".embeddings.position_embeddings.weight") if g else None,
                        gtype=st.g(e + ".embeddings.token_type_embeddings.weight") if g else None)
        # Conv3d visual front-end: packed bf16 weights, fp32 biases / cls / learned positions
        v = e + ".encoder.visn_fc"
        self.fe = dict(w1=st.w(v + ".conv.1.weight"), b1=st.p(v + ".conv.1.bias"),
                       w2=st.w(v + ".conv.4.weight"), b2=st.p(v + ".conv.4.bias"),
                       cls=st.p(v + ".cls_token").view(-1), pe=st.p(v + ".position_encoding.pe.weight"),
                       gw1=st.packed(st.grad, v + ".conv.1.weight") if g else None,
                       gb1=st.g(v + ".conv.1.bias") if g else None,
                       gw2=st.packed(st.grad, v + ".conv.4.weight") if g else None,
                       gb2=st.g(v + ".conv.4.bias") if g else None,
                       gcls=st.g(v + ".cls_token").view(-1) if g else None,
                       gpe=st.g(v + ".position_encoding.pe.weight") if g else None)
        self.l_layers = [self._bert_layer("%s.encoder.layer.%d" % (e, i)) for i in range(self.nl)]
        self.r_layers = [self._bert_layer("%s.encoder.r_layers.%d" % (e, i)) for i in range(self.nr)]
        self.x_layer = self._cross_layer(e + ".encoder.cross_attn_layer.cross")
        self.pooler = self._lin(e + ".pooler_dict.cross.dense2.weight")
        self.logit_fc = self._head("logit_fc")
        if self.task != "hgqa":
            return
        self.hgx_layer = self._cross_layer("hgq_encoder.cross_attn_layer.cross")
        self.hg_pooler = self._lin("hgq_encoder.pooler_dict.cross.dense2.weight")
        self.rel_q = self._hg_embed("relation_query_embed")
        self.act_q = self._hg_embed("action_query_embed")
        self.rel_dec = [self._decoder_layer("rel_decoder.layers.%d" % i) for i in range(self.nd)]
        self.act_dec = [self._decoder_layer("action_decoder.layers.%d" % i) for i in range(self.nd)]
        self.class_embed = self._head("class_embed")
        self.action_embed = self._head("action_embed")
        # hoisted memory K/V projection of all decoder layers: packed bf16 [n_dec*2*768, 768] + fp32 bias
        n = 2 * self.nd
        self.kv_w = torch.empty(n * 2 * H, H, device=self.dev, dtype=BF16)
        self.kv_b = torch.empty(n * 2 * H, device=self.dev, dtype=torch.float32)
        self.tok = torch.empty((self.n_act + self.n_rel + 1) * H, device=self.dev, dtype=BF16)

    def _hg_embed(self, p):
        st = self.st
        g = st.grad is not None
        return dict(word=st.p(p + ".word_embeddings.weight"), type=st.p(p + ".token_type_embeddings.weight"),
                    ln=self._norm(p + ".LayerNorm"),
                    gword=st.g(p + ".word_embeddings.weight") if g else None,
                    gtype=st.g(p + ".token_type_embeddings.weight") if g else None)

    def _dec_layers(self):
        return self.rel_dec + self.act_dec

    def refresh_derived(self):
        """bf16 copies that are not plain shadows: the packed K/V weights of the 10 decoder layers and the
        CrossEncoder's type / cls tokens.  Call after the weights change (load or optimiser step)."""
        if self.task != "hgqa":
            return
        st = self.st
        for i, name in enumerate(["rel_decoder.layers.%d" % j for j in range(self.nd)]
                                 + ["action_decoder.layers.%d" % j for j in range(self.nd)]):
            w = st.p(name + ".multihead_attn.in_proj_weight")
            ops.to_bf16(w[H:], out=self.kv_w[i * 2 * H:(i + 1) * 2 * H])
            self.kv_b[i * 2 * H:(i + 1) * 2 * H].copy_(st.p(name + ".multihead_attn.in_proj_bias")[H:])
        # [act_token x n_act | rel_token x n_rel | cls_token]
        for j in range(self.n_act):
            ops.to_bf16(st.p("hgq_encoder.act_token").view(-1), out=self.tok[j * H:(j + 1) * H])
        for j in range(self.n_rel):
            o = (self.n_act + j) * H
            ops.to_bf16(st.p("hgq_encoder.rel_token").view(-1), out=self.tok[o:o + H])
        o = (self.n_act + self.n_rel) * H
        ops.to_bf16(st.p("hgq_encoder.cls_token").view(-1), out=self.tok[o:o + H])

    def _drop(self, p, active=True):
        """(p, fresh call-site salt, seed) for a dropout site of this forward, or None.  Sites are live in every
        train-mode forward — also in the forward-only LXRT cross layers, as in the reference's model.train()."""
        if not (active and self._training and self.dropout_on and p > 0):
            return None
        self._salt += 1
        return (p, self._salt, self.drop_seed)

    # ------------------------------------------------------------------ Conv3d visual front-end
    def frontend_fwd(self, feat, save):
        """VisualFeatEncoder (lxrt/modeling_capsbert.py:989-996, 1037-1073): feat fp32 [B,2048,T,7,7] as the backbone
        emits it (or bf16 channels-last [B,T,7,7,2048]) -> visual tokens bf16 [B*(1+(T-8)*49), 768].  Two implicit-GEMM
        launches; the second one's epilogue writes GELU + learned position + dropout straight into the token matrix
        behind each clip's cls row."""
        F = self.fe
        x0 = ops.ncdhw_to_ndhwc(feat) if feat.dtype == torch.float32 else feat
        B, T = x0.shape[0], x0.shape[1]
        a1, pre1 = ops.conv3d_fwd(x0, F["w1"], F["b1"], preact=True if save else None)
        a1 = a1.view(B, T - 4, 7, 7, H)
        nv = 1 + (T - 8) * 49
        assert nv <= F["pe"].shape[0], "clip too long for the learned position table"
        tokens = torch.empty(B * nv, H, device=self.dev, dtype=BF16)
        dv = self._drop(self.p_hidden)
        _, pre2 = ops.conv3d_fwd(a1, F["w2"], F["b2"], out=tokens, rows_per_clip=nv, row_off=1,
                                 preact=True if save else None, pos_emb=F["pe"], drop=dv)
        ops.frontend_cls(F["cls"], F["pe"], tokens, B, nv, drop=dv)
        return tokens, (x0, pre1, a1, pre2, dv, nv)

    def frontend_bwd(self, saved, d_tok, need_dfeat=False):
        """Backward of frontend_fwd: position/cls gradients, both weight gradients (packed layout, straight into the
        flat gradient buffer), bias gradients, and on request the gradient wrt the backbone features (fp32 NCDHW)."""
        F = self.fe
        x0, pre1, a1, pre2, dv, nv = saved
        B, T = x0.shape[0], x0.shape[1]
        d_pre2 = ops.frontend_bwd(d_tok, pre2, F["gpe"], F["gcls"], B, nv, drop=dv)
        d_pre2v = d_pre2.view(B, T - 8, 7, 7, H)
        self._conv_wgrad(d_pre2v, a1, F["gw2"], F["gb2"])
        d_pre1 = ops.conv3d_dgrad(d_pre2v, F["w2"], H, preact=pre1)
        self._conv_wgrad(d_pre1, x0, F["gw1"], F["gb1"])
        if not need_dfeat:
            return None
        return ops.ndhwc_to_ncdhw(ops.conv3d_dgrad(d_pre1, F["w1"], x0.shape[-1]))

    def _conv_wgrad(self, dy, x, gw, gb):
        if self.use_streams:
            self.sw.wait_event(torch.cuda.current_stream().record_event())
            self._keep.append((dy, x))
        with self._on(self.sw):
            ops.conv3d_wgrad(dy, x, gw)
            ops.colsum_accum(dy.view(-1, dy.shape[-1]), gb)

    # ------------------------------------------------------------------ primitive fwd/bwd helpers
    def _wgrad(self, lin, dy, x, bias=True):
        """dW += dy^T x (split-K fp32 atomics into the flat gradient buffer), db += colsum(dy) — on the wgrad
        stream, ordered after whatever produced dy on the current stream."""
        if not self.use_streams:
            ops.gemm(dy, x, a_t=True, b_t=True, out=lin.gw, accumulate=True)
            if bias:
                ops.colsum_accum(dy, lin.gb)
            return
        self.sw.wait_event(torch.cuda.current_stream().record_event())
        self._keep.append((dy, x))
        with self._on(self.sw):
            ops.gemm(dy, x, a_t=True, b_t=True, out=lin.gw, accumulate=True)
            if bias:
                ops.colsum_accum(dy, lin.gb)

    def _bias_grad(self, dy, gb):
        if not self.use_streams:
            ops.colsum_accum(dy, gb)
            return
        self.sw.wait_event(torch.cuda.current_stream().record_event())
        self._keep.append((dy,))
        with self._on(self.sw):
            ops.colsum_accum(dy, gb)

    def _ffn_fwd(self, f1, f2, ln, a, save, act="gelu", eps=EPS_BERT, addend=None, p_out=0.0, p_inner=0.0):
        pre = torch.empty(a.shape[0], f1.w.shape[0], device=self.dev, dtype=BF16) if (save and act == "gelu") else None
        d_in, d_out = self._drop(p_inner), self._drop(p_out)
        h = ops.gemm(a, f1.w, bias=f1.b, act=act, preact=pre, drop=d_in)
        s2 = ops.gemm(h, f2.w, bias=f2.b, residual=a, drop=d_out)
        y, y2, m, r = ops.layernorm_fwd(s2, ln.g, ln.b, eps, addend=addend, save_stats=save)
        return y, y2, (a, pre, h, s2, m, r, d_in, d_out)

    def _ln_bwd_split(self, dy, x, ln, m, r, drop, dy2=None, dxsum=None):
        """LayerNorm backward of x = dropout(dense) + residual → (gradient for the residual branch, gradient for
        the projection); identical tensors when the site had no dropout."""
        res = ops.layernorm_bwd(dy, x, ln.g, m, r, ln.gg, ln.gb, dy2=dy2, dxsum=dxsum, out_drop=drop)
        return res if drop is not None else (res, res)

    def _ffn_bwd(self, f1, f2, ln, saved, dy, dy2=None, act="gelu"):
        a, pre, h, s2, m, r, d_in, d_out = saved
        ds2, ds2d = self._ln_bwd_split(dy, s2, ln, m, r, d_out, dy2=dy2, dxsum=f2.gb)
        self._wgrad(f2, ds2d, h, bias=False)
        if act == "gelu":
            dpre = ops.gemm(ds2d, f2.w, b_t=True, mul_aux=pre, mul_mode="gelu_grad")
        else:
            # h went through ReLU and (decoders) dropout: h > 0 marks "active and kept"
            dpre = ops.gemm(ds2d, f2.w, b_t=True, mul_aux=h, mul_mode="relu_mask",
                            mul_scale=1.0 / (1.0 - d_in[0]) if d_in is not None else 1.0)
        self._wgrad(f1, dpre, a)
        return ops.gemm(dpre, f1.w, b_t=True, residual=ds2)      # d a  (FFN path + residual path)

    # ------------------------------------------------------------------ BertLayer (L, R)
    def _bert_fwd(self, L, x, B, n, key_mask, save):
        qkv = ops.gemm(x, L["qkv"].w, bias=L["qkv"].b)
        da = self._drop(self.p_attn)
        ctx, lse = ops.attention_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, n, n, key_mask=key_mask,
                                     save_lse=save, drop=da)
        d1 = self._drop(self.p_hidden)
        s1 = ops.gemm(ctx, L["o"].w, bias=L["o"].b, residual=x, drop=d1)
        a, _, m1, r1 = ops.layernorm_fwd(s1, L["ln1"].g, L["ln1"].b, EPS_BERT, save_stats=save)
        y, _, ffn = self._ffn_fwd(L["f1"], L["f2"], L["ln2"], a, save, p_out=self.p_hidden)
        return y, (x, qkv, ctx, lse, s1, m1, r1, ffn, d1, da)

    def _bert_bwd(self, L, saved, dy, B, n, key_mask, need_dx=True):
        x, qkv, ctx, lse, s1, m1, r1, ffn, d1, d_att = saved
        da = self._ffn_bwd(L["f1"], L["f2"], L["ln2"], ffn, dy)
        ds1, ds1d = self._ln_bwd_split(da, s1, L["ln1"], m1, r1, d1, dxsum=L["o"].gb)
        self._wgrad(L["o"], ds1d, ctx, bias=False)
        dctx = ops.gemm(ds1d, L["o"].w, b_t=True)
        dqkv = torch.empty_like(qkv)
        ops.attention_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], ctx, dctx, lse, dqkv[:, :H], dqkv[:, H:2 * H],
                          dqkv[:, 2 * H:], B, n, n, key_mask=key_mask, drop=d_att)
        self._wgrad(L["qkv"], dqkv, x)
        if not need_dx:
            return None
        return ops.gemm(dqkv, L["qkv"].w, b_t=True, residual=ds1)

    # ------------------------------------------------------------------ CrossLayer (X, HG)
    def _cross_fwd(self, X, lang, visn, B, nl, nv, lang_kmask, visn_kmask, save):
        q = X["qkv"]
        qkv_l = ops.gemm(lang, q.w, bias=q.b)
        qkv_v = ops.gemm(visn, q.w, bias=q.b)
        # one attention block, both directions (lxrt/modeling_capsbert.py:643-644)
        dal, dav = self._drop(self.p_attn), self._drop(self.p_attn)
        ctx_l, lse_l = ops.attention_fwd(qkv_l[:, :H], qkv_v[:, H:2 * H], qkv_v[:, 2 * H:], B, nl, nv,
                                         key_mask=visn_kmask, save_lse=save, drop=dal)
        ctx_v, lse_v = ops.attention_fwd(qkv_v[:, :H], qkv_l[:, H:2 * H], qkv_l[:, 2 * H:], B, nv, nl,
                                         key_mask=lang_kmask, save_lse=save, drop=dav)
        dl, dv_ = self._drop(self.p_hidden), self._drop(self.p_hidden)
        s_l = ops.gemm(ctx_l, X["o"].w, bias=X["o"].b, residual=lang, drop=dl)
        s_v = ops.gemm(ctx_v, X["o"].w, bias=X["o"].b, residual=visn, drop=dv_)
        a_l, _, ml, rl = ops.layernorm_fwd(s_l, X["lno"].g, X["lno"].b, EPS_BERT, save_stats=save)
        a_v, _, mv, rv = ops.layernorm_fwd(s_v, X["lno"].g, X["lno"].b, EPS_BERT, save_stats=save)
        y_l, _, ffn_l = self._ffn_fwd(X["lf1"], X["lf2"], X["lln"], a_l, save, p_out=self.p_hidden)
        y_v, _, ffn_v = self._ffn_fwd(X["vf1"], X["vf2"], X["vln"], a_v, save, p_out=self.p_hidden)
        return y_l, y_v, (lang, visn, qkv_l, qkv_v, ctx_l, lse_l, ctx_v, lse_v, s_l, ml, rl, s_v, mv, rv, ffn_l, ffn_v,
                          dl, dv_, dal, dav)

    def _cross_bwd(self, X, saved, dy_l, dy_v, B, nl, nv, lang_kmask, visn_kmask):
        (lang, visn, qkv_l, qkv_v, ctx_l, lse_l, ctx_v, lse_v, s_l, ml, rl, s_v, mv, rv, ffn_l, ffn_v,
         dl, dv_, dal, dav) = saved
        da_l = self._ffn_bwd(X["lf1"], X["lf2"], X["lln"], ffn_l, dy_l)
        da_v = self._ffn_bwd(X["vf1"], X["vf2"], X["vln"], ffn_v, dy_v)
        ds_l, ds_ld = self._ln_bwd_split(da_l, s_l, X["lno"], ml, rl, dl, dxsum=X["o"].gb)
        ds_v, ds_vd = self._ln_bwd_split(da_v, s_v, X["lno"], mv, rv, dv_, dxsum=X["o"].gb)
        self._wgrad(X["o"], ds_ld, ctx_l, bias=False)
        self._wgrad(X["o"], ds_vd, ctx_v, bias=False)
        dctx_l = ops.gemm(ds_ld, X["o"].w, b_t=True)
        dctx_v = ops.gemm(ds_vd, X["o"].w, b_t=True)
        dqkv_l, dqkv_v = torch.empty_like(qkv_l), torch.empty_like(qkv_v)
        ops.attention_bwd(qkv_l[:, :H], qkv_v[:, H:2 * H], qkv_v[:, 2 * H:], ctx_l, dctx_l, lse_l,
                          dqkv_l[:, :H], dqkv_v[:, H:2 * H], dqkv_v[:, 2 * H:], B, nl, nv, key_mask=visn_kmask,
                          drop=dal)
        ops.attention_bwd(qkv_v[:, :H], qkv_l[:, H:2 * H], qkv_l[:, 2 * H:], ctx_v, dctx_v, lse_v,
                          dqkv_v[:, :H], dqkv_l[:, H:2 * H], dqkv_l[:, 2 * H:], B, nv, nl, key_mask=lang_kmask,
                          drop=dav)
        self._wgrad(X["qkv"], dqkv_l, lang)
        self._wgrad(X["qkv"], dqkv_v, visn)
        d_lang = ops.gemm(dqkv_l, X["qkv"].w, b_t=True, residual=ds_l)
        d_visn = ops.gemm(dqkv_v, X["qkv"].w, b_t=True, residual=ds_v)
        return d_lang, d_visn

    # ------------------------------------------------------------------ DETR decoder layer (D)
    def _dec_fwd(self, D, tgt, tgt_p, qpos, k_mem, v_mem, B, nq, per, save, first):
        qk = ops.gemm(tgt_p, D["sa_qk"].w, bias=D["sa_qk"].b)
        v = ops.gemm(tgt, D["sa_v"].w, bias=D["sa_v"].b)
        da1, da2 = self._drop(self.p_dec), self._drop(self.p_dec)
        sa, lse1 = ops.attention_fwd(qk[:, :H], qk[:, H:], v, B, nq, nq, frames=(per, per), save_lse=save, drop=da1)
        d1, d2 = self._drop(self.p_dec), self._drop(self.p_dec)
        s1 = ops.gemm(sa, D["sa_o"].w, bias=D["sa_o"].b, residual=tgt, drop=d1)
        t1, t1p, m1, r1 = ops.layernorm_fwd(s1, D["n1"].g, D["n1"].b, EPS_DETR, addend=qpos, save_stats=save)
        cq = ops.gemm(t1p, D["ca_q"].w, bias=D["ca_q"].b)
        ca, lse2 = ops.attention_fwd(cq, k_mem, v_mem, B, nq, self.nv_mem, save_lse=save, drop=da2)
        s2 = ops.gemm(ca, D["ca_o"].w, bias=D["ca_o"].b, residual=t1, drop=d2)
        t2, _, m2, r2 = ops.layernorm_fwd(s2, D["n2"].g, D["n2"].b, EPS_DETR, save_stats=save)
        t3, t3p, ffn = self._ffn_fwd(D["l1"], D["l2"], D["n3"], t2, save, act="relu", eps=EPS_DETR, addend=qpos,
                                     p_out=self.p_dec, p_inner=self.p_dec)
        return t3, t3p, (tgt, tgt_p, qk, v, sa, lse1, s1, m1, r1, t1p, cq, ca, lse2, s2, m2, r2, ffn, first, d1, d2,
                         da1, da2)

    def _dec_bwd(self, D, saved, dt3, dt3p, dqpos, k_mem, v_mem, dk_mem, dv_mem, B, nq, per):
        (tgt, tgt_p, qk, v, sa, lse1, s1, m1, r1, t1p, cq, ca, lse2, s2, m2, r2, ffn, first, d1, d2,
         da1, da2) = saved
        if dt3p is not None:
            self._acc(dqpos, dt3p)
        dt2 = self._ffn_bwd(D["l1"], D["l2"], D["n3"], ffn, dt3, dy2=dt3p, act="relu")
        ds2, ds2d = self._ln_bwd_split(dt2, s2, D["n2"], m2, r2, d2, dxsum=D["ca_o"].gb)
        self._wgrad(D["ca_o"], ds2d, ca, bias=False)
        dca = ops.gemm(ds2d, D["ca_o"].w, b_t=True)
        dcq = torch.empty_like(cq)
        ops.attention_bwd(cq, k_mem, v_mem, ca, dca, lse2, dcq, dk_mem, dv_mem, B, nq, self.nv_mem, drop=da2)
        self._wgrad(D["ca_q"], dcq, t1p)
        dt1p = ops.gemm(dcq, D["ca_q"].w, b_t=True)
        self._acc(dqpos, dt1p)
        ds1, ds1d = self._ln_bwd_split(ds2, s1, D["n1"], m1, r1, d1, dy2=dt1p, dxsum=D["sa_o"].gb)
        self._wgrad(D["sa_o"], ds1d, sa, bias=False)
        dsa = ops.gemm(ds1d, D["sa_o"].w, b_t=True)
        dqk, dv = torch.empty_like(qk), torch.empty_like(v)
        ops.attention_bwd(qk[:, :H], qk[:, H:], v, sa, dsa, lse1, dqk[:, :H], dqk[:, H:], dv, B, nq, nq,
                          frames=(per, per), drop=da1)
        self._wgrad(D["sa_qk"], dqk, tgt_p)
        dtgt_p = ops.gemm(dqk, D["sa_qk"].w, b_t=True)
        if first:
            # tgt is the constant zero tensor (tasks/agqa_model.py:235): only the bias of the value projection learns
            self._bias_grad(dv, D["sa_v"].gb)
            self._acc(dqpos, dtgt_p)
            return None, None
        self._wgrad(D["sa_v"], dv, tgt)
        dtgt = ops.gemm(dv, D["sa_v"].w, b_t=True, residual=ds1)
        return dtgt, dtgt_p

    # ------------------------------------------------------------------ both DETR decoders in lock-step
    # The relation and the action decoder have the same structure and depth and different weights.  Run one after the
    # other — or on two streams — their 5-30 us kernels on 4096 / 1536 rows mostly serialise (persistent one-CTA-per-SM
    # GEMMs).  Here their token rows are stacked ([B*128 relation rows | B*48 action rows]) and every projection /
    # LayerNorm of a layer position is ONE grouped launch (rows >= split read the action decoder's weights); only the
    # attention calls (different query counts) and the weight-gradient products stay per decoder.
    def _g2(self, x, A, Bm, split, **kw):
        return ops.gemm(x, A.w, bias=A.b, b2=Bm.w, bias2=Bm.b, split_m=split, **kw)

    def _g2t(self, dy, A, Bm, split, **kw):
        return ops.gemm(dy, A.w, b_t=True, b2=Bm.w, split_m=split, **kw)

    def _wgrad2(self, A, Bm, dy, x, split, bias=True):
        """Weight (and bias) gradients of the same layer of both decoders: one split-K launch whose K rows below the
        split accumulate into the relation decoder's gradient and the rest into the action decoder's."""
        if split % 64 or not self.use_streams:
            self._wgrad(A, dy[:split], x[:split], bias=bias)
            self._wgrad(Bm, dy[split:], x[split:], bias=bias)
            return
        self.sw.wait_event(torch.cuda.current_stream().record_event())
        self._keep.append((dy, x))
        with self._on(self.sw):
            ops.gemm(dy, x, a_t=True, b_t=True, out=A.gw, accumulate=True, d2=Bm.gw, k_split=split)
            if bias:
                ops.colsum_accum(dy[:split], A.gb)
                ops.colsum_accum(dy[split:], Bm.gb)

    def _ln2_fwd(self, x, NA, NB, split, eps, addend=None, save=False):
        return ops.layernorm_fwd(x, NA.g, NA.b, eps, addend=addend, save_stats=save, group=(split, NB.g, NB.b))

    def _ln2_bwd(self, dy, x, NA, NB, split, m, r, drop, dy2=None, dxsum=None):
        res = ops.layernorm_bwd(dy, x, NA.g, m, r, NA.gg, NA.gb, dy2=dy2, dxsum=dxsum[0] if dxsum else None,
                                out_drop=drop, group=(split, NB.g, NB.gg, NB.gb, dxsum[1] if dxsum else None))
        return res if drop is not None else (res, res)

    def _attn2_fwd(self, q, k, v, B, nq, lk, split, frames, save, drop, kv_rel=None, kv_act=None):
        """Self-attention (k, v stacked like q) or cross-attention over the visual memory (kv_rel / kv_act =
        (k, v) column slices of the hoisted K/V buffer) for both decoders; two launches into one stacked output."""
        out = torch.empty(q.shape[0], H, device=self.dev, dtype=BF16)
        res = []
        self._fork(self.s2)                     # the action decoder's (smaller) problem runs beside the relation one
        for g, (r0, r1) in enumerate(((0, split), (split, q.shape[0]))):
            n = nq[g]
            if kv_rel is None:
                kk, vv, lkk, fr = k[r0:r1], v[r0:r1], n, (frames[g], frames[g])
            else:
                kk, vv = (kv_rel, kv_act)[g]
                lkk, fr = lk, None
            with self._on(self.s2 if g == 1 else None):
                _, sv = ops.attention_fwd(q[r0:r1], kk, vv, B, n, lkk, frames=fr, save_lse=save, drop=drop[g],
                                          out=out[r0:r1])
            res.append(sv)
        self._join(self.s2)
        return out, res

    def _dec2_fwd(self, DA, DB, tgt, tgt_p, qpos, kv_rel, kv_act, B, nq, per, split, save, first):
        qk = self._g2(tgt_p, DA["sa_qk"], DB["sa_qk"], split)
        v = self._g2(tgt, DA["sa_v"], DB["sa_v"], split)
        da1 = (self._drop(self.p_dec), self._drop(self.p_dec))
        da2 = (self._drop(self.p_dec), self._drop(self.p_dec))
        sa, lse1 = self._attn2_fwd(qk[:, :H], qk[:, H:], v, B, nq, None, split, per, save, da1)
        d1, d2 = self._drop(self.p_dec), self._drop(self.p_dec)
        s1 = self._g2(sa, DA["sa_o"], DB["sa_o"], split, residual=tgt, drop=d1)
        t1, t1p, m1, r1 = self._ln2_fwd(s1, DA["n1"], DB["n1"], split, EPS_DETR, addend=qpos, save=save)
        cq = self._g2(t1p, DA["ca_q"], DB["ca_q"], split)
        ca, lse2 = self._attn2_fwd(cq, None, None, B, nq, self.nv_mem, split, None, save, da2, kv_rel=kv_rel, kv_act=kv_act)
        s2 = self._g2(ca, DA["ca_o"], DB["ca_o"], split, residual=t1, drop=d2)
        t2, _, m2, r2 = self._ln2_fwd(s2, DA["n2"], DB["n2"], split, EPS_DETR, save=save)
        d_in, d_out = self._drop(self.p_dec), self._drop(self.p_dec)
        h = self._g2(t2, DA["l1"], DB["l1"], split, act="relu", drop=d_in)
        s3 = self._g2(h, DA["l2"], DB["l2"], split, residual=t2, drop=d_out)
        t3, t3p, m3, r3 = self._ln2_fwd(s3, DA["n3"], DB["n3"], split, EPS_DETR, addend=qpos, save=save)
        return t3, t3p, (tgt, tgt_p, qk, v, sa, lse1, s1, m1, r1, t1p, cq, ca, lse2, s2, m2, r2, t2, h, s3, m3, r3,
                         d_in, d_out, first, d1, d2, da1, da2)

    def _dec2_bwd(self, DA, DB, saved, dt3, dt3p, dqpos, kv_rel, kv_act, dkv_rel, dkv_act, B, nq, per, split):
        (tgt, tgt_p, qk, v, sa, lse1, s1, m1, r1, t1p, cq, ca, lse2, s2, m2, r2, t2, h, s3, m3, r3, d_in, d_out, first,
         d1, d2, da1, da2) = saved
        rng = ((0, split), (split, tgt.shape[0]))
        if dt3p is not None:
            self._acc(dqpos, dt3p)
        # FFN
        ds3, ds3d = self._ln2_bwd(dt3, s3, DA["n3"], DB["n3"], split, m3, r3, d_out, dy2=dt3p,
                                  dxsum=(DA["l2"].gb, DB["l2"].gb))
        self._wgrad2(DA["l2"], DB["l2"], ds3d, h, split, bias=False)
        dpre = self._g2t(ds3d, DA["l2"], DB["l2"], split, mul_aux=h, mul_mode="relu_mask",
                         mul_scale=1.0 / (1.0 - d_in[0]) if d_in is not None else 1.0)
        self._wgrad2(DA["l1"], DB["l1"], dpre, t2, split)
        dt2 = self._g2t(dpre, DA["l1"], DB["l1"], split, residual=ds3)
        # cross-attention
        ds2, ds2d = self._ln2_bwd(dt2, s2, DA["n2"], DB["n2"], split, m2, r2, d2, dxsum=(DA["ca_o"].gb, DB["ca_o"].gb))
        self._wgrad2(DA["ca_o"], DB["ca_o"], ds2d, ca, split, bias=False)
        dca = self._g2t(ds2d, DA["ca_o"], DB["ca_o"], split)
        dcq = torch.empty_like(cq)
        self._fork(self.s2)
        for g, (a0, a1) in enumerate(rng):
            kk, vv = (kv_rel, kv_act)[g]
            dkk, dvv = (dkv_rel, dkv_act)[g]
            with self._on(self.s2 if g == 1 else None):
                ops.attention_bwd(cq[a0:a1], kk, vv, ca[a0:a1], dca[a0:a1], lse2[g], dcq[a0:a1], dkk, dvv, B, nq[g],
                                  self.nv_mem, drop=da2[g])
        self._join(self.s2)
        self._wgrad2(DA["ca_q"], DB["ca_q"], dcq, t1p, split)
        dt1p = self._g2t(dcq, DA["ca_q"], DB["ca_q"], split)
        self._acc(dqpos, dt1p)
        # self-attention
        ds1, ds1d = self._ln2_bwd(ds2, s1, DA["n1"], DB["n1"], split, m1, r1, d1, dy2=dt1p,
                                  dxsum=(DA["sa_o"].gb, DB["sa_o"].gb))
        self._wgrad2(DA["sa_o"], DB["sa_o"], ds1d, sa, split, bias=False)
        dsa = self._g2t(ds1d, DA["sa_o"], DB["sa_o"], split)
        dqk, dv = torch.empty_like(qk), torch.empty_like(v)
        self._fork(self.s2)
        for g, (a0, a1) in enumerate(rng):
            with self._on(self.s2 if g == 1 else None):
                ops.attention_bwd(qk[a0:a1, :H], qk[a0:a1, H:], v[a0:a1], sa[a0:a1], dsa[a0:a1], lse1[g],
                                  dqk[a0:a1, :H], dqk[a0:a1, H:], dv[a0:a1], B, nq[g], nq[g],
                                  frames=(per[g], per[g]), drop=da1[g])
        self._join(self.s2)
        self._wgrad2(DA["sa_qk"], DB["sa_qk"], dqk, tgt_p, split)
        dtgt_p = self._g2t(dqk, DA["sa_qk"], DB["sa_qk"], split)
        if first:
            # tgt is the constant zero tensor (tasks/agqa_model.py:235): only the bias of the value projection learns
            self._bias_grad(dv[:split], DA["sa_v"].gb)
            self._bias_grad(dv[split:], DB["sa_v"].gb)
            self._acc(dqpos, dtgt_p)
            return None, None
        self._wgrad2(DA["sa_v"], DB["sa_v"], dv, tgt, split)
        dtgt = self._g2t(dv, DA["sa_v"], DB["sa_v"], split, residual=ds1)
        return dtgt, dtgt_p

    def _decoders2_fwd(self, qpos, kv_mem, B, nq, per, split, save):
        tgt = torch.zeros_like(qpos)
        tgt_p = qpos
        saved = []
        for i in range(self.nd):
            cr, ca = i * 2 * H, (self.nd + i) * 2 * H
            tgt, tgt_p, s = self._dec2_fwd(self.rel_dec[i], self.act_dec[i], tgt, tgt_p, qpos,
                                           (kv_mem[:, cr:cr + H], kv_mem[:, cr + H:cr + 2 * H]),
                                           (kv_mem[:, ca:ca + H], kv_mem[:, ca + H:ca + 2 * H]), B, nq, per, split, save,
                                           first=(i == 0))
            saved.append(s)
        return tgt, saved

    def _decoders2_bwd(self, saved, d_out, kv_mem, dkv_mem, B, nq, per, split):
        dqpos = torch.zeros_like(d_out)
        dt3, dt3p = d_out, None
        for i in reversed(range(self.nd)):
            cr, ca = i * 2 * H, (self.nd + i) * 2 * H
            sl = lambda t, c: (t[:, c:c + H], t[:, c + H:c + 2 * H])
            dt3, dt3p = self._dec2_bwd(self.rel_dec[i], self.act_dec[i], saved[i], dt3, dt3p, dqpos, sl(kv_mem, cr),
                                       sl(kv_mem, ca), sl(dkv_mem, cr), sl(dkv_mem, ca), B, nq, per, split)
        return dqpos

    @staticmethod
    def _acc(dst, src):
        ops.copy3d(src, dst, 1, 1, src.numel(), 0, 0, 0, 0, accumulate=True)

    def _decoder_fwd(self, layers, qpos, kv_mem, layer0, B, nq, per, save):
        tgt = torch.zeros_like(qpos)
        tgt_p = qpos
        saved = []
        for i, D in enumerate(layers):
            c = (layer0 + i) * 2 * H
            tgt, tgt_p, s = self._dec_fwd(D, tgt, tgt_p, qpos, kv_mem[:, c:c + H], kv_mem[:, c + H:c + 2 * H], B, nq,
                                          per, save, first=(i == 0))
            saved.append(s)
        return tgt, saved

    def _decoder_bwd(self, layers, saved, d_out, kv_mem, dkv_mem, layer0, B, nq, per):
        dqpos = torch.zeros_like(d_out)
        dt3, dt3p = d_out, None
        for i in reversed(range(len(layers))):
            c = (layer0 + i) * 2 * H
            dt3, dt3p = self._dec_bwd(layers[i], saved[i], dt3, dt3p, dqpos, kv_mem[:, c:c + H],
                                      kv_mem[:, c + H:c + 2 * H], dkv_mem[:, c:c + H], dkv_mem[:, c + H:c + 2 * H],
                                      B, nq, per)
        return dqpos

    # ------------------------------------------------------------------ heads, poolers, embeddings
    def _head_fwd(self, Hd, x, save):
        pre = torch.empty(x.shape[0], 2 * H, device=self.dev, dtype=BF16) if save else None
        h1 = ops.gemm(x, Hd["l0"].w, bias=Hd["l0"].b, act="gelu", preact=pre)
        n, _, m, r = ops.layernorm_fwd(h1, Hd["ln"].g, Hd["ln"].b, EPS_BERT, save_stats=save)
        c = Hd["l3"].w.shape[0]
        ld = (c + 3) // 4 * 4
        logits = torch.empty(x.shape[0], ld, device=self.dev, dtype=torch.float32)[:, :c]
        ops.gemm(n, Hd["l3"].w, bias=Hd["l3"].b, out=logits)
        return logits, (x, pre, h1, m, r, n)

    def _head_bwd(self, Hd, saved, dlogits, need_dx=True):
        x, pre, h1, m, r, n = saved
        self._wgrad(Hd["l3"], dlogits, n)
        dn = ops.gemm(dlogits, Hd["l3"].w, b_t=True)
        dh1 = ops.layernorm_bwd(dn, h1, Hd["ln"].g, m, r, Hd["ln"].gg, Hd["ln"].gb)
        dpre = ops.act_bwd(dh1, pre, "gelu")
        self._wgrad(Hd["l0"], dpre, x)
        return ops.gemm(dpre, Hd["l0"].w, b_t=True) if need_dx else None

    def _pool_fwd(self, P, first, n_first, second, n_second, B):
        """BertPooler2 (lxrt/modeling_capsbert.py:1525-1533): tanh(W [first[:,0] ; second[:,0]] + b)."""
        cat = torch.empty(B, 2 * H, device=self.dev, dtype=BF16)
        ops.copy3d(first, cat, B, 1, H, n_first * H, 0, 2 * H, 0)
        ops.copy3d(second, cat[:, H:], B, 1, H, n_second * H, 0, 2 * H, 0)
        pooled = ops.gemm(cat, P.w, bias=P.b, act="tanh")
        return pooled, (cat, pooled)

    def _pool_bwd(self, P, saved, dpooled, d_first, n_first, d_second, n_second, B):
        cat, pooled = saved
        dz = ops.act_bwd(dpooled, pooled, "tanh")
        self._wgrad(P, dz, cat)
        dcat = ops.gemm(dz, P.w, b_t=True)
        ops.copy3d(dcat, d_first, B, 1, H, 2 * H, 0, n_first * H, 0, accumulate=True)
        ops.copy3d(dcat[:, H:], d_second, B, 1, H, 2 * H, 0, n_second * H, 0, accumulate=True)

    def _hg_embed_fwd(self, E, seg_ids, B, nq, save, p_drop=0.0):
        rows = B * nq
        y = torch.empty(rows, H, device=self.dev, dtype=BF16)
        pre = torch.empty_like(y) if save else None
        mean = torch.empty(rows, device=self.dev, dtype=torch.float32) if save else None
        rstd = torch.empty_like(mean) if save else None
        d = self._drop(p_drop)
        ops.call("shg_embed_ln_fwd", ops.ptr(E["word"]), None, None, ops.ptr(E["type"]), ops.ptr(seg_ids),
                 ops.ptr(E["ln"].g), ops.ptr(E["ln"].b), EPS_BERT, ops.ptr(y), ops.ptr(pre), ops.ptr(mean),
                 ops.ptr(rstd), rows, nq, H, d[0] if d else 0.0, d[1] if d else 0, ops.ptr(d[2]) if d else None,
                 ops.stream())
        return y, (seg_ids, pre, mean, rstd, d)

    def _hg_embed_bwd(self, E, saved, dq, B, nq):
        seg_ids, pre, mean, rstd, d = saved
        de = ops.layernorm_bwd(dq, pre, E["ln"].g, mean, rstd, E["ln"].gg, E["ln"].gb, in_drop=d)
        # word table is used whole (no lookup → no padding mask); frame-id table is nn.Embedding(padding_idx=0)
        ops.call("shg_embed_scatter_bwd", ops.ptr(de), None, ops.ptr(seg_ids), ops.ptr(E["gword"]), -1, None, -1,
                 ops.ptr(E["gtype"]), 0, B * nq, nq, H, ops.stream())

    # ------------------------------------------------------------------ whole path
    def forward(self, inp, train=False):
        """inp (device tensors): either visn_tokens bf16 [B*393,768] (output of a stock Conv3d front-end + cls + pos)
        or feat (backbone features, fp32 [B,2048,16,7,7] / bf16 channels-last) for the built-in front-end;
        input_ids/input_mask/segment_ids int64 [B,40], rel_segment_ids [B,128], act_segment_ids [B,48]."""
        save = train
        # LXRT cross layers / pooler / first answer head: forward-only in --taskHGQA training (the loss never reads
        # `logit`) unless the caller wants their gradient (--taskVQA training, tasks/agqaVQA.py)
        save_x = train and (self.task != "hgqa" or bool(inp.get("x_grad", False)))
        B = inp["input_ids"].shape[0]
        ctx = dict(B=B)
        ids, seg = inp["input_ids"].reshape(-1), inp["segment_ids"].reshape(-1)
        nl = inp["input_ids"].shape[1]
        lang = torch.empty(B * nl, H, device=self.dev, dtype=BF16)
        pre = torch.empty_like(lang) if save else None
        mean = torch.empty(B * nl, device=self.dev, dtype=torch.float32) if save else None
        rstd = torch.empty_like(mean) if save else None
        E = self.emb
        self._salt, self._training = 0, bool(train)
        d_emb = self._drop(self.p_emb)
        ops.call("shg_embed_ln_fwd", ops.ptr(E["word"]), ops.ptr(ids), ops.ptr(E["pos"]), ops.ptr(E["type"]),
                 ops.ptr(seg), ops.ptr(E["ln"].g), ops.ptr(E["ln"].b), EPS_BERT, ops.ptr(lang), ops.ptr(pre),
                 ops.ptr(mean), ops.ptr(rstd), B * nl, nl, H, d_emb[0] if d_emb else 0.0, d_emb[1] if d_emb else 0,
                 ops.ptr(d_emb[2]) if d_emb else None, ops.stream())
        ctx["emb"] = (ids, seg, pre, mean, rstd, nl, d_emb)
        # additive key mask (1-mask)*-10000 (lxrt/modeling_capsbert.py:1826-1834); the visual mask is all zeros (:1836)
        kmask = ((1.0 - inp["input_mask"].to(torch.float32)) * -10000.0).contiguous()
        ctx["kmask"] = kmask
        # language layers on a side stream, relation layers on the main one
        self._fork(self.s1)
        ctx["L"] = []
        with self._on(self.s1):
            for L in self.l_layers:
                lang, s = self._bert_fwd(L, lang, B, nl, kmask, save)
                ctx["L"].append(s)
        lang_l = lang
        if "visn_tokens" in inp:
            visn = inp["visn_tokens"]
        else:
            visn, ctx["frontend"] = self.frontend_fwd(inp["feat"], save)
        nv = visn.shape[0] // B
        assert visn.shape[0] == B * nv and nv <= 416, "visual memory: %d rows for %d clips" % (visn.shape[0], B)
        self.nv_mem = nv                 # memory length of the decoders' cross-attention (393 for 16-frame clips)
        ctx["R"] = []
        for L in self.r_layers:
            visn, s = self._bert_fwd(L, visn, B, nv, None, save)
            ctx["R"].append(s)
        visn_r = visn
        ctx["visn_r"] = visn_r
        self._join(self.s1)
        out = dict(lang_l=lang_l, visn_r=visn_r, kmask=kmask)
        hgqa = self.task == "hgqa"
        if hgqa:
            # hoisted memory K/V projection for all decoder layers (tasks/agqa_model.py:218-249)
            kv_mem = ops.gemm(visn_r, self.kv_w, bias=self.kv_b)
            ctx["kv_mem"] = kv_mem
            self._fork(self.s1, self.s2)
        if inp.get("skip_x", False) is False:
            # LXRT cross layers + pooler + answer head: forward-only in --taskHGQA training → side stream
            with self._on(self.s1) if hgqa else self._on(None):
                xl, xv = lang_l, visn_r
                ctx["X"] = []
                for _ in range(self.nx):             # the same module twice (lxrt/modeling_capsbert.py:1247-1249)
                    xl, xv, sx = self._cross_fwd(self.x_layer, xl, xv, B, nl, nv, kmask, None, save_x)
                    ctx["X"].append(sx)
                pooled, ctx["x_pool"] = self._pool_fwd(self.pooler, xv, nv, xl, nl, B)
                out["logit"], ctx["x_head"] = self._head_fwd(self.logit_fc, pooled, save_x)
                out["lang"], out["visn"], out["pooled"] = xl, xv, pooled
        ctx["save_x"] = save_x
        ctx["dims"] = (nl, nv, 0, 0, 0)
        if not hgqa:
            return out, ctx
        nq_r, nq_a = self.T * self.n_rel, self.T * self.n_act
        if self.fused_decoders:
            # both decoders in lock-step over stacked token rows (see _dec2_fwd)
            split = B * nq_r
            rel_q, ctx["rel_q"] = self._hg_embed_fwd(self.rel_q, inp["rel_segment_ids"].reshape(-1), B, nq_r, save,
                                                     self.p_relq)
            act_q, ctx["act_q"] = self._hg_embed_fwd(self.act_q, inp["act_segment_ids"].reshape(-1), B, nq_a, save,
                                                     self.p_actq)
            qpos = torch.cat([rel_q, act_q], dim=0)
            dec_out, ctx["dec2"] = self._decoders2_fwd(qpos, kv_mem, B, (nq_r, nq_a), (self.n_rel, self.n_act), split,
                                                       save)
            rel_out, act_out = dec_out[:split], dec_out[split:]
            self._fork(self.s2)
            out["rel_logits"], ctx["rel_head"] = self._head_fwd(self.class_embed, rel_out, save)
            with self._on(self.s2):
                out["act_logits"], ctx["act_head"] = self._head_fwd(self.action_embed, act_out, save)
            self._join(self.s1, self.s2)
        else:
            rel_q, ctx["rel_q"] = self._hg_embed_fwd(self.rel_q, inp["rel_segment_ids"].reshape(-1), B, nq_r, save,
                                                     self.p_relq)
            rel_out, ctx["rel_dec"] = self._decoder_fwd(self.rel_dec, rel_q, kv_mem, 0, B, nq_r, self.n_rel, save)
            out["rel_logits"], ctx["rel_head"] = self._head_fwd(self.class_embed, rel_out, save)
            with self._on(self.s2):
                act_q, ctx["act_q"] = self._hg_embed_fwd(self.act_q, inp["act_segment_ids"].reshape(-1), B, nq_a, save,
                                                         self.p_actq)
                act_out, ctx["act_dec"] = self._decoder_fwd(self.act_dec, act_q, kv_mem, self.nd, B, nq_a, self.n_act,
                                                            save)
                out["act_logits"], ctx["act_head"] = self._head_fwd(self.action_embed, act_out, save)
            self._join(self.s1, self.s2)
        out["rel_out"], out["act_out"] = rel_out, act_out
        # hypergraph tokens: [cls | per frame: actions + act_token, relations + rel_token]
        per = self.n_act + self.n_rel
        ng = self.T * per + 1
        hg = torch.empty(B * ng, H, device=self.dev, dtype=BF16)
        tok = self.tok
        ops.copy3d(tok[per * H:], hg, B, 1, H, 0, 0, ng * H, 0)
        ops.copy3d(act_out, hg[1:], B, self.T, self.n_act * H, nq_a * H, self.n_act * H, ng * H, per * H,
                   add=tok[:self.n_act * H])
        ops.copy3d(rel_out, hg[1 + self.n_act:], B, self.T, self.n_rel * H, nq_r * H, self.n_rel * H, ng * H, per * H,
                   add=tok[self.n_act * H:per * H])
        hg_kmask = None
        if inp.get("hg_key_mask") is not None:     # --useHGMask (tasks/agqa_model.py:261-264); [B, 177] additive
            hg_kmask = inp["hg_key_mask"]
        ctx["hg_kmask"] = hg_kmask
        xl, xg = lang_l, hg
        ctx["HGX"] = []
        for _ in range(self.nx):
            xl, xg, s = self._cross_fwd(self.hgx_layer, xl, xg, B, nl, ng, kmask, hg_kmask, save)
            ctx["HGX"].append(s)
        hg_pooled, ctx["hg_pool"] = self._pool_fwd(self.hg_pooler, xg, ng, xl, nl, B)
        out["hg_logit"], ctx["hg_head"] = self._head_fwd(self.logit_fc, hg_pooled, save)
        out["hg_pooled"] = hg_pooled
        ctx["dims"] = (nl, nv, ng, nq_r, nq_a)
        return out, ctx

    def backward(self, ctx, d_hg_logit, d_rel_logits, d_act_logits, need_dvisn=False, d_logit=None):
        """Gradients of the --taskHGQA loss (hg answer loss + relation CE + action CE, tasks/agqaHGQA.py:344-378)
        accumulated into the flat gradient buffer.  d_* are bf16 [rows, n_cls] views with ld % 8 == 0.
        d_logit: gradient of the LXRT answer logits (`logit`) — the --taskVQA loss (tasks/agqaVQA.py); needs a
        forward that saved the cross layers (task 'vqa', or inp['x_grad'])."""
        if d_logit is not None:
            self.backward_x(ctx, d_logit)
        if self.task == "hgqa":
            self.backward_late_params(ctx, d_hg_logit, d_rel_logits, d_act_logits)
        return self.backward_early_params(ctx, need_dvisn)

    def backward_x(self, ctx, d_logit):
        """Backward of the LXRT answer branch: logit_fc -> BertPooler2 -> the shared CrossLayer applied twice
        (lxrt/modeling_capsbert.py:658-677, 1247-1249, 1525-1533).  Leaves the gradients w.r.t. the language-layer
        and relation-layer outputs in ctx['d_x_lang'] / ctx['d_x_visn'] for backward_early_params."""
        if not ctx.get("save_x"):
            raise RuntimeError("the LXRT cross layers were run forward-only: pass x_grad=True (or use task='vqa') to "
                               "backpropagate through `logit`")
        B = ctx["B"]
        nl, nv = ctx["dims"][0], ctx["dims"][1]
        kmask = ctx["kmask"]
        d_pool = self._head_bwd(self.logit_fc, ctx["x_head"], d_logit)
        d_xv = torch.zeros(B * nv, H, device=self.dev, dtype=BF16)
        d_xl = torch.zeros(B * nl, H, device=self.dev, dtype=BF16)
        self._pool_bwd(self.pooler, ctx["x_pool"], d_pool, d_xv, nv, d_xl, nl, B)
        for sx in reversed(ctx["X"]):
            d_xl, d_xv = self._cross_bwd(self.x_layer, sx, d_xl, d_xv, B, nl, nv, kmask, None)
        ctx["d_x_lang"], ctx["d_x_visn"] = d_xl, d_xv

    def backward_late_params(self, ctx, d_hg_logit, d_rel_logits, d_act_logits):
        """First half of the backward: answer head, CrossEncoder, classifier heads, both decoders, query embeddings
        and the hoisted K/V weights.  On return (all side streams joined) the gradients of every parameter that
        sits AFTER the LXRT encoder in the flat layout are final — the data-parallel trainer starts their
        all-reduce here, overlapped with backward_early_params."""
        B = ctx["B"]
        nl, nv, ng, nq_r, nq_a = ctx["dims"]
        self.nv_mem = nv
        kmask = ctx["kmask"]
        per = self.n_act + self.n_rel
        # answer head → pooler → CrossEncoder
        d_pool = self._head_bwd(self.logit_fc, ctx["hg_head"], d_hg_logit)
        d_xg = torch.zeros(B * ng, H, device=self.dev, dtype=BF16)
        d_xl = torch.zeros(B * nl, H, device=self.dev, dtype=BF16)
        self._pool_bwd(self.hg_pooler, ctx["hg_pool"], d_pool, d_xg, ng, d_xl, nl, B)
        for s in reversed(ctx["HGX"]):
            d_xl, d_xg = self._cross_bwd(self.hgx_layer, s, d_xl, d_xg, B, nl, ng, kmask, ctx["hg_kmask"])
        ctx["d_lang_l"] = d_xl
        if self.early_lang_bwd:
            self.backward_lang_start(ctx)
        # hypergraph assembly backward: cls / type-token gradients and the decoder-output gradients
        st = self.st
        ops.colsum_accum(d_xg.view(B, ng * H)[:, :H], st.g("hgq_encoder.cls_token").view(-1))
        d_act_hg = torch.empty(B * nq_a, H, device=self.dev, dtype=BF16)
        d_rel_hg = torch.empty(B * nq_r, H, device=self.dev, dtype=BF16)
        ops.copy3d(d_xg[1:], d_act_hg, B, self.T, self.n_act * H, ng * H, per * H, nq_a * H, self.n_act * H)
        ops.copy3d(d_xg[1 + self.n_act:], d_rel_hg, B, self.T, self.n_rel * H, ng * H, per * H, nq_r * H,
                   self.n_rel * H)
        ops.colsum_accum(d_act_hg, st.g("hgq_encoder.act_token").view(-1))
        ops.colsum_accum(d_rel_hg, st.g("hgq_encoder.rel_token").view(-1))
        kv_mem = ctx["kv_mem"]
        dkv_mem = torch.empty_like(kv_mem)
        ctx["dkv_mem"] = dkv_mem
        if "dec2" in ctx:
            # both decoders in lock-step (see _dec2_bwd): head backward per decoder into one stacked gradient
            split = B * nq_r
            d_out = torch.empty(B * (nq_r + nq_a), H, device=self.dev, dtype=BF16)
            self._fork(self.s2)
            with self._on(self.s2):
                d_act_out = self._head_bwd(self.action_embed, ctx["act_head"], d_act_logits)
                ops.copy3d(d_act_out, d_out[split:], 1, 1, d_act_out.numel(), 0, 0, 0, 0)
            d_rel_out = self._head_bwd(self.class_embed, ctx["rel_head"], d_rel_logits)
            ops.copy3d(d_rel_out, d_out[:split], 1, 1, d_rel_out.numel(), 0, 0, 0, 0)
            self._join(self.s2)
            self._acc(d_out[:split], d_rel_hg)
            self._acc(d_out[split:], d_act_hg)
            dq = self._decoders2_bwd(ctx["dec2"], d_out, kv_mem, dkv_mem, B, (nq_r, nq_a), (self.n_rel, self.n_act),
                                     split)
            self._hg_embed_bwd(self.rel_q, ctx["rel_q"], dq[:split], B, nq_r)
            self._hg_embed_bwd(self.act_q, ctx["act_q"], dq[split:], B, nq_a)
        else:
            self._fork(self.s2)
            # side stream 2: action head + decoder; main: relation head + decoder (disjoint parameters and dK/dV slices)
            with self._on(self.s2):
                d_act_out = self._head_bwd(self.action_embed, ctx["act_head"], d_act_logits)
                self._acc(d_act_out, d_act_hg)
                dq_a = self._decoder_bwd(self.act_dec, ctx["act_dec"], d_act_out, kv_mem, dkv_mem, self.nd, B, nq_a,
                                         self.n_act)
                self._hg_embed_bwd(self.act_q, ctx["act_q"], dq_a, B, nq_a)
            d_rel_out = self._head_bwd(self.class_embed, ctx["rel_head"], d_rel_logits)
            self._acc(d_rel_out, d_rel_hg)
            dq_r = self._decoder_bwd(self.rel_dec, ctx["rel_dec"], d_rel_out, kv_mem, dkv_mem, 0, B, nq_r, self.n_rel)
            self._hg_embed_bwd(self.rel_q, ctx["rel_q"], dq_r, B, nq_r)
            self._join(self.s2)
        # hoisted K/V projection: one wgrad per decoder layer (their parameters belong to this half)
        mem = ctx["visn_r"]
        for i, D in enumerate(self._dec_layers()):
            c = i * 2 * H
            self._wgrad(D["ca_kv"], dkv_mem[:, c:c + 2 * H], mem)
        # (the weight-gradient stream is NOT drained here: a data-parallel trainer issues the all-reduce of these
        # parameters from that stream, see HotPathTrainer._reduce_async; a single-GPU step lets it run on)

    # The second half of the backward in steps, so that a data-parallel trainer can start the all-reduce of each
    # group of gradients the moment it is final (HotPathTrainer.train_step): begin -> relation layers nr-1 .. 0 ->
    # end.  backward_early_params() below runs them in sequence (optionally in two parts).
    def backward_lang_start(self, ctx):
        """Language layers + embeddings on side stream s1.  Their input gradient (CrossEncoder branch + LXRT answer
        branch) is complete as soon as the CrossEncoder's backward has run, i.e. BEFORE the decoders' backward: started
        there, the 40-token language side fills the SMs that the decoders' chains of small kernels leave idle instead
        of competing with the relation layers' large kernels later.  Idempotent per step."""
        if ctx.get("_lang_started"):
            return
        ctx["_lang_started"] = True
        B = ctx["B"]
        nl = ctx["dims"][0]
        kmask = ctx["kmask"]
        # gradient w.r.t. the language-layer output: CrossEncoder branch (+ LXRT answer branch)
        d_lang, dxl = ctx.get("d_lang_l"), ctx.get("d_x_lang")
        if d_lang is None:
            d_lang = dxl
        elif dxl is not None:
            self._acc(d_lang, dxl)
        self._fork(self.s1)
        with self._on(self.s1):
            for i in reversed(range(self.nl)):
                d_lang = self._bert_bwd(self.l_layers[i], ctx["L"][i], d_lang, B, nl, kmask)
            ids, seg, pre, mean, rstd, _, d_emb = ctx["emb"]
            E = self.emb
            de = ops.layernorm_bwd(d_lang, pre, E["ln"].g, mean, rstd, E["ln"].gg, E["ln"].gb, in_drop=d_emb)
            ops.call("shg_embed_scatter_bwd", ops.ptr(de), ops.ptr(ids), ops.ptr(seg), ops.ptr(E["gword"]), 0,
                     ops.ptr(E["gpos"]), 0, ops.ptr(E["gtype"]), 0, B * nl, nl, H, ops.stream())
        ctx["_lang_pending"] = True

    def backward_early_begin(self, ctx):
        """The language side (if backward_late_params has not started it already) and the hoisted K/V dgrad on the main
        stream, which leaves the gradient w.r.t. the relation-layer output."""
        self.backward_lang_start(ctx)
        if "dkv_mem" in ctx:
            ctx["d_visn_mid"] = ops.gemm(ctx["dkv_mem"], self.kv_w, b_t=True, residual=ctx.get("d_x_visn"))
        else:
            ctx["d_visn_mid"] = ctx["d_x_visn"]

    def backward_r_layer(self, ctx, i, need_dvisn=False):
        B, nv = ctx["B"], ctx["dims"][1]
        need0 = need_dvisn or "frontend" in ctx
        ctx["d_visn_mid"] = self._bert_bwd(self.r_layers[i], ctx["R"][i], ctx["d_visn_mid"], B, nv, None,
                                           need_dx=(i > 0 or need0))

    def backward_lang_join(self, ctx):
        """The main stream waits for the language side: its weight gradients are then queued on the wgrad stream."""
        if ctx.pop("_lang_pending", False):
            self._join(self.s1)

    def backward_early_end(self, ctx, need_dvisn=False):
        if "frontend" in ctx:
            ctx["d_feat"] = self.frontend_bwd(ctx["frontend"], ctx["d_visn_mid"], need_dfeat=need_dvisn)
        self.backward_lang_join(ctx)
        self._join(self.sw)
        self._keep.clear()
        return ctx["d_visn_mid"]

    def backward_early_params(self, ctx, need_dvisn=False, part=None, split=2, join=True):
        """Second half: the hoisted K/V dgrad, relation layers (main stream), language layers and embeddings
        (side stream).  part=None runs all of it; part=0 runs the whole language side and the relation layers with
        index >= split, part=1 the remaining relation layers and the Conv3d front-end."""
        first, second = part in (None, 0), part in (None, 1)
        if first:
            self.backward_early_begin(ctx)
            r_range = range(self.nr - 1, (split if part == 0 else 0) - 1, -1)
        else:
            r_range = range(min(split, self.nr) - 1, -1, -1)
        for i in r_range:
            self.backward_r_layer(ctx, i, need_dvisn)
        if first:
            self.backward_lang_join(ctx)
        if second:
            return self.backward_early_end(ctx, need_dvisn)
        if join:
            self._join(self.sw)
        return ctx["d_visn_mid"]
