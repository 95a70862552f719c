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
