"""Multi-rank host logic on CPU (gloo, world_size 2): the one collective of the path — the all-reduce of the
contiguous live-gradient span — and the per-rank sharding of the synthetic batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from shg_vqa_b200.params import ParamStore, build_key_shapes
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import all_reduce_live_grads, live_in_hgqa_training, shard_seed
    ks = build_key_shapes("vqa")            # smaller key set is enough for the host logic
    st = ParamStore(ks, "cpu", live=live_in_hgqa_training)
    g = torch.Generator().manual_seed(100 + rank)
    st.grad.copy_(torch.randn(st.n_total, generator=g))
    before = st.grad.clone()
    all_reduce_live_grads(st, world)
    # reference: sum of both ranks' buffers over the live span, untouched elsewhere
    other = torch.randn(st.n_total, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    live = st.live_extent
    ok_live = torch.allclose(st.grad[live], before[live] + other[live], atol=1e-6)
    frozen = torch.ones(st.n_total, dtype=torch.bool)
    frozen[live] = False
    ok_frozen = torch.equal(st.grad[frozen], before[frozen])
    # sharding: ranks draw different clips, same shapes
    a = make_inputs(2, seed=shard_seed(1234, rank), with_feat=False, visn_tokens=True)
    ids_sum = int(a["input_ids"].sum())
    q.put((rank, ok_live, ok_frozen, ids_sum, st.n_live, st.live_begin))
    dist.destroy_process_group()


def test_live_gradient_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, live0, frozen0, ids0, n0, b0), (r1, live1, frozen1, ids1, n1, b1) = res
    assert live0 and live1, "live span must hold the sum over ranks"
    assert frozen0 and frozen1, "tensors that never receive gradients must not be exchanged"
    assert ids0 != ids1, "ranks must draw different shards"
    assert (n0, b0) == (n1, b1)


def test_data_parallel_groups_tile_the_live_span():
    """The gradient groups of the overlapped data-parallel schedule (ParamStore.dp_groups, reduced by
    HotPathTrainer.train_step as each becomes final) must cover the live gradient span exactly once — with and without
    the Conv3d front-end — or a key-order change would silently leave gradients unreduced / reduced twice.  Also pins
    what belongs where: nothing of the LXRT encoder is 'late', every relation layer is its own group."""
    import torch
    from shg_vqa_b200.params import ParamStore, build_key_shapes
    from shg_vqa_b200.trainer import live_in_hgqa_training, live_with_frontend
    ks = build_key_shapes("hgqa")
    for live in (live_in_hgqa_training, live_with_frontend):
        st = ParamStore(ks, "cpu", with_grad=False, live=live)
        cover = torch.zeros(st.n_total, dtype=torch.int32)
        for name, slices in st.dp_groups.items():
            for sl in slices:
                cover[sl] += 1
        assert int(cover[st.live_extent].min()) == 1 and int(cover[st.live_extent].max()) == 1
        assert int(cover.sum()) == st.n_live
        want = {"late", "lang", "emb", "vec", "r0", "r1", "r2", "r3", "r4"} | ({"front"} if live is live_with_frontend else set())
        assert set(st.dp_groups) == want

        def group_of(key):
            off = st.offsets[key][0]
            return next(n for n, sls in st.dp_groups.items() if any(s.start <= off < s.stop for s in sls))
        e = "lxrt_encoder.model.bert.encoder."
        assert group_of("rel_decoder.layers.0.linear1.weight") == "late"
        assert group_of("hgq_encoder.cross_attn_layer.cross.lang_inter.dense.weight") == "late"
        assert group_of("class_embed.3.weight") == "late" and group_of("class_embed.3.bias") == "vec"
        for i in range(5):
            assert group_of(e + "layer.%d.intermediate.dense.weight" % i) == "lang"
            assert group_of(e + "r_layers.%d.output.dense.weight" % i) == "r%d" % i
            assert group_of(e + "r_layers.%d.attention.self.query.weight" % i) == "r%d" % i
            assert group_of(e + "r_layers.%d.output.LayerNorm.weight" % i) == "vec"
        assert group_of("lxrt_encoder.model.bert.embeddings.word_embeddings.weight") == "emb"
        if live is live_with_frontend:
            assert group_of(e + "visn_fc.conv.1.weight") == "front"
        for k in st.live_keys:                      # no live tensor of the LXRT encoder may be reduced early
            if k.startswith("lxrt_encoder."):
                assert group_of(k) != "late", k
