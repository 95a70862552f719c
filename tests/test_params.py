"""The state_dict boundary: key names, order, shapes, dtypes and aliasing equal the reference's
(golden lists produced from the unmodified reference by tools/gen_golden.py)."""
import json
import os

import torch

from shg_vqa_b200 import params as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_hgqa_key_layout_equals_reference():
    want = _golden("state_dict_keys.json")
    got = P.build_key_shapes("hgqa")
    assert list(got) == list(want)
    for k, (shape, dtype) in got.items():
        assert list(shape) == want[k][0] and dtype == want[k][1], k
    assert len(got) == 669


def test_vqa_key_layout_equals_reference():
    want = _golden("state_dict_keys_vqa.json")
    got = P.build_key_shapes("vqa")
    assert list(got) == list(want)
    assert all(list(got[k][0]) == want[k][0] for k in got)
    assert len(got) == 324


def test_store_roundtrip_aliases_and_fused_views():
    ks = P.build_key_shapes("hgqa")
    st = P.ParamStore(ks, "cpu")
    from oracle import shg_oracle as so
    sd = so.make_state_dict({k: [list(s), d] for k, (s, d) in ks.items()}, seed=3)
    st.load_state_dict({"module." + k: v for k, v in sd.items()}, strict=True)     # DataParallel prefix accepted
    out = st.state_dict()
    assert list(out) == list(ks)
    for k in ks:
        assert torch.equal(out[k], sd[k]), k
    e = "lxrt_encoder.model.bert.encoder."
    a = out[e + "x_layers.0.lang_inter.dense.weight"]
    assert a.data_ptr() == out[e + "x_layers.1.lang_inter.dense.weight"].data_ptr() \
        == out[e + "cross_attn_layer.cross.lang_inter.dense.weight"].data_ptr()
    # q|k|v adjacent → one [2304,768] view
    p = e + "layer.0.attention.self."
    qkv = st.fused(st.master, p + "query.weight", 2304, 768)
    assert torch.equal(qkv, torch.cat([sd[p + "query.weight"], sd[p + "key.weight"], sd[p + "value.weight"]]))
    b = st.fused(st.master, p + "query.bias", 2304)
    assert torch.equal(b, torch.cat([sd[p + "query.bias"], sd[p + "key.bias"], sd[p + "value.bias"]]))
    # strict loading refuses a missing key
    bad = dict(sd); bad.pop("logit_fc.3.bias")
    try:
        st.load_state_dict(bad, strict=True)
        assert False
    except RuntimeError:
        pass


def test_conv_weights_are_stored_packed_and_exposed_in_reference_layout(key_shapes):
    """The Conv3d weights of the visual front-end live packed [c_out, 5, 3, 3, c_in] in the flat buffers (what the
    implicit-GEMM kernels read); state_dict()/load_state_dict() keep the reference's [c_out, c_in, 5, 3, 3]."""
    import torch
    from shg_vqa_b200.params import ParamStore, build_key_shapes
    from shg_vqa_b200.trainer import live_with_frontend, live_in_hgqa_training
    ks = build_key_shapes("hgqa")
    st = ParamStore(ks, "cpu", live=live_with_frontend)
    k = "lxrt_encoder.model.bert.encoder.visn_fc.conv.4.weight"
    w = torch.randn(768, 768, 5, 3, 3)
    st.p(k).copy_(w)
    assert tuple(st.state_dict()[k].shape) == (768, 768, 5, 3, 3)
    assert torch.equal(st.state_dict()[k], w)
    packed = st.packed(st.master, k)
    assert packed.shape == (768, 45 * 768) and packed.is_contiguous()
    assert torch.equal(packed, w.permute(0, 2, 3, 4, 1).reshape(768, -1))
    # live with the front-end, frozen without; either way inside the bf16-shadowed matrix region
    off, n, _ = st.offsets[k]
    assert st.live_begin <= off and off + n <= st.live_end and off + n <= st.n_matrix
    st2 = ParamStore(ks, "cpu", live=live_in_hgqa_training)
    off2, n2, _ = st2.offsets[k]
    assert off2 + n2 <= st2.live_begin and off2 + n2 <= st2.n_matrix
    for dead in ("box_fc.weight", "pos_layer_norm.weight"):
        assert not live_with_frontend("lxrt_encoder.model.bert.encoder.visn_fc." + dead)
    for live in ("conv.1.bias", "cls_token", "position_encoding.pe.weight"):
        assert live_with_frontend("lxrt_encoder.model.bert.encoder.visn_fc." + live)


def test_data_parallel_buckets_tile_the_live_span():
    """late | early-upper | early-rest (the three all-reduce phases of the captured data-parallel step) cover the
    live gradient span exactly once, with and without the front-end's parameters."""
    from shg_vqa_b200.params import ParamStore, build_key_shapes
    from shg_vqa_b200.trainer import live_in_hgqa_training, live_with_frontend
    for live in (live_in_hgqa_training, live_with_frontend):
        st = ParamStore(build_key_shapes("hgqa"), "cpu", live=live)
        spans = sorted((s.start, s.stop) for s in st.bucket_early_upper + st.bucket_early_rest + [st.bucket_late])
        pos = st.live_begin
        for a, b in spans:
            assert a == pos and b > a and a % 8 == 0
            pos = b
        assert pos == st.live_end
        # the upper bucket: every language-layer matrix, relation layers 2..4, the BertEmbeddings tables
        n_upper = sum(s.stop - s.start for s in st.bucket_early_upper)
        per_layer = 4 * 768 * 768 + 2 * 768 * 3072
        n_emb = (30522 + 512 + 2 + 2) * 768
        assert n_upper == (5 + 3) * per_layer + n_emb
        in_upper = lambda k: any(s.start <= st.offsets[k][0] < s.stop for s in st.bucket_early_upper)
        e = "lxrt_encoder.model.bert."
        assert in_upper(e + "encoder.r_layers.3.output.dense.weight")
        assert in_upper(e + "encoder.layer.0.output.dense.weight")
        assert in_upper(e + "embeddings.word_embeddings.weight") and in_upper(e + "embeddings.LayerNorm.bias")
        assert not in_upper(e + "encoder.r_layers.1.output.dense.weight")
        assert not in_upper(e + "encoder.layer.0.output.dense.bias")


def test_param_tree_mirrors_reference_module_names(key_shapes):
    """compat.build_param_tree: the nn.Module view of the flat store has exactly the reference's state_dict keys in
    the reference's order, its 576 de-duplicated parameters, aliases sharing one Parameter, and strict loading
    writes through to the flat master buffer (what the engine and the fused optimiser read)."""
    import torch
    from shg_vqa_b200.compat import _Node, build_param_tree
    from shg_vqa_b200.params import ParamStore, build_key_shapes, canonical_key
    ks = build_key_shapes("hgqa", 171, 457, 158)
    st = ParamStore(ks, "cpu", with_grad=True)
    root = _Node()
    params = build_param_tree(root, st, ks)
    sd = root.state_dict()
    assert list(sd) == list(key_shapes)
    assert len(list(root.parameters())) == 576 == len(params)
    named = dict(root.named_parameters(remove_duplicate=False))
    e = "lxrt_encoder.model.bert.encoder."
    assert named[e + "x_layers.0.visn_inter.dense.weight"] is named[e + "cross_attn_layer.cross.visn_inter.dense.weight"]
    assert named[e + "x_layers.1.visn_inter.dense.weight"] is named[e + "cross_attn_layer.cross.visn_inter.dense.weight"]
    assert hasattr(root.lxrt_encoder, "load") and root.lxrt_encoder.dim == 768
    g = torch.Generator().manual_seed(3)
    new = {}
    for k, v in sd.items():
        ck = canonical_key(k)
        if ck not in new:
            new[ck] = torch.randn(v.shape, generator=g) if v.is_floating_point() else v.clone()
        new[k] = new[ck]
    root.load_state_dict({"module." + k if False else k: v for k, v in new.items()}, strict=True)
    for k in ("logit_fc.3.weight", e + "visn_fc.conv.4.weight", "rel_decoder.layers.2.self_attn.in_proj_weight"):
        assert torch.equal(st.p(k), new[k]), k
    # a gradient view attached to a Parameter aliases the flat gradient buffer
    p = named["logit_fc.3.bias"]
    p.grad = st.g("logit_fc.3.bias")
    p.grad.fill_(2.0)
    off, n, _ = st.offsets["logit_fc.3.bias"]
    assert float(st.grad[off:off + n].sum()) == 2.0 * n


def test_local_slow_r50_backbone_surface():
    """The stock-PyTorch backbone outside the hot path: pytorchvideo's module names (so hub / reference checkpoints
    load), 31.6 M parameters, res5 feature map with the temporal length preserved and 1/32 of the spatial size."""
    import torch
    from shg_vqa_b200.backbone import VideoBackbone
    vb = VideoBackbone("slow_r50", freeze=True)
    keys = list(vb.state_dict())
    assert keys[0] == "backbone.blocks.0.conv.weight"
    assert "backbone.blocks.3.res_blocks.5.branch2.conv_a.weight" in keys
    assert vb.state_dict()["backbone.blocks.3.res_blocks.0.branch2.conv_a.weight"].shape == (256, 512, 3, 1, 1)
    assert vb.state_dict()["backbone.blocks.1.res_blocks.0.branch2.conv_a.weight"].shape == (64, 64, 1, 1, 1)
    assert abs(sum(p.numel() for p in vb.parameters()) - 31_634_496) == 0
    assert isinstance(vb.backbone.blocks[-1], torch.nn.Identity)          # video_encoder.py:45
    y = vb.encode(torch.randn(1, 3, 4, 64, 64))
    assert tuple(y.shape) == (1, 2048, 4, 2, 2) and y.dtype == torch.float32 and not y.requires_grad
