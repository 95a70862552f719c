"""The captured (CUDA-graph) train step and the host-buffer entry point against the eager step.

Same weights, same inputs, dropout off: the replayed graph, the host path with direct H2D copies and the host path
with the next batch prefetched on the copy stream must reproduce the eager step.  The first step's losses (a pure
forward) agree to 1e-5; fp32 atomics in the split-K weight gradients make gradient sums order-dependent, and BertAdam
without bias correction turns a sign flip of a noise-level gradient into a +-3.2*lr update, so after the first
update >= 99.9 % of the weights must agree to 1e-6 and the later steps' losses to 5e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(dev, seed):
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import csr_from_padded
    raw = make_inputs(4, seed=seed, with_feat=False, visn_tokens=True)
    inp = {k: raw[k] for k in ("input_ids", "input_mask", "segment_ids", "rel_segment_ids", "act_segment_ids")}
    inp["visn_tokens"] = raw["visn_tokens"].reshape(-1, 768).to(torch.bfloat16)
    batch = {"target": raw["target"]}
    batch["rel_labels"], batch["rel_off"] = csr_from_padded(raw["rel_tokens"], raw["rel_lengths"])
    batch["act_labels"], batch["act_off"] = csr_from_padded(raw["act_tokens"], raw["act_lengths"])
    # CSR label vectors padded to a fixed capacity so that every batch fits the captured buffers
    for fam, cap in (("rel", 4 * 16 * 8), ("act", 4 * 16 * 3)):
        lab = torch.zeros(cap, dtype=torch.int32)
        lab[: batch[fam + "_labels"].numel()] = batch[fam + "_labels"]
        batch[fam + "_labels"] = lab
    pin = lambda d: {k: v.contiguous().pin_memory() for k, v in d.items()}
    return pin(inp), pin(batch)


def _trainer(dev):
    from shg_vqa_b200.trainer import HotPathTrainer
    tr = HotPathTrainer(dev, lr=1e-3, t_total=-1, dropout=False)
    tr.init_weights(seed=11)
    return tr


def test_graph_and_host_steps_match_eager():
    dev = torch.device("cuda:0")
    a = _make(dev, 1)
    b = _make(dev, 2)
    to_dev = lambda d: {k: v.to(dev) for k, v in d.items()}

    ref = _trainer(dev)
    losses_ref = []
    w_ref = None
    for h_inp, h_batch in (a, b, a):
        _, scal = ref.train_step(to_dev(h_inp), to_dev(h_batch))
        losses_ref.append(ref.losses_from_scalars(scal))
        if w_ref is None:
            w_ref = ref.store.master.clone()

    for mode in ("graph", "host", "host_prefetch", "host_lagged"):
        tr = _trainer(dev)
        w0 = tr.store.master.clone()
        tr.capture(to_dev(a[0]), to_dev(a[1]), warmup=1)
        # capture() warms up with real optimiser steps but leaves NO trace of them: weights, Adam moments, the
        # bf16 shadow, gradients and the step count are what they were (the trajectory must not drift)
        assert torch.equal(tr.store.master, w0) and tr.step_count == 0
        assert float(tr.exp_avg.abs().max()) == 0.0 and float(tr.exp_avg_sq.abs().max()) == 0.0
        assert float(tr.store.grad.abs().max()) == 0.0
        assert torch.equal(tr.store.shadow, w0[:tr.store.n_matrix].to(torch.bfloat16))
        seq = (a, b, a)
        got = []
        w1 = None
        for i, (h_inp, h_batch) in enumerate(seq):
            if mode == "graph":
                _, scal = tr.train_step_graph(to_dev(h_inp), to_dev(h_batch))
                got.append(tr.losses_from_scalars(scal))
            elif mode == "host":
                got.append(tr.train_step_host(h_inp, h_batch))
            elif mode == "host_prefetch":
                nxt = seq[i + 1] if i + 1 < len(seq) else (None, None)
                got.append(tr.train_step_host(h_inp, h_batch, next_inp=nxt[0], next_batch=nxt[1]))
            else:
                # prefetch + lagged loss read: call i hands out the loss of call i-1
                nxt = seq[i + 1] if i + 1 < len(seq) else (None, None)
                prev = tr.train_step_host(h_inp, h_batch, next_inp=nxt[0], next_batch=nxt[1], lagged=True)
                assert (prev is None) == (i == 0)
                if prev is not None:
                    got.append(prev)
                if i + 1 == len(seq):
                    got.append(tr.drain_losses())
                    assert tr.drain_losses() is None
            if w1 is None:
                torch.cuda.synchronize()
                w1 = tr.store.master.clone()
        for i, (g, r) in enumerate(zip(got, losses_ref)):
            tol = 1e-5 if i == 0 else 5e-3
            for k in ("qa", "rel_ce", "act_ce", "total"):
                assert abs(g[k] - r[k]) <= tol * abs(r[k]) + 1e-6, (mode, i, k, g[k], r[k])
        live = tr.store.live_extent
        agree = float(((w1[live] - w_ref[live]).abs() <= 1e-6).float().mean())
        assert agree >= 0.999, (mode, agree)
        assert float((w1[live] - w_ref[live]).abs().max()) <= 2.1 * 3.17e-3, mode      # at worst a sign flip


def test_split_second_backward_half_equals_unsplit():
    """backward_early_params in two parts (the data-parallel schedule reduces the upper layers' gradients between
    them) accumulates the same gradients as the single pass."""
    dev = torch.device("cuda:0")
    h_inp, h_batch = _make(dev, 3)
    inp = {k: v.to(dev) for k, v in h_inp.items()}
    batch = {k: v.to(dev) for k, v in h_batch.items()}
    grads = []
    for split_run in (False, True):
        tr = _trainer(dev)
        out, ctx = tr.engine.forward(inp, train=True)
        scal, d_hg, d_rel, d_act, _ = tr.loss_and_grads(out, batch)
        tr.engine.backward_late_params(ctx, d_hg, d_rel, d_act)
        if split_run:
            tr.engine.backward_early_params(ctx, part=0, join=False)     # as the data-parallel schedule calls it
            tr.engine.backward_early_params(ctx, part=1)
        else:
            tr.engine.backward_early_params(ctx)
        torch.cuda.synchronize()
        grads.append(tr.store.grad[tr.store.live_extent].clone())
    a, b = grads
    assert float((a - b).norm() / a.norm()) <= 1e-5
    for sl in tr.store.bucket_early_upper + tr.store.bucket_early_rest:
        ga, gb = tr.store.grad.new_zeros(0), None
        lo, hi = sl.start - tr.store.live_begin, sl.stop - tr.store.live_begin
        assert float(a[lo:hi].abs().max()) > 0                      # every bucket actually receives gradients
        assert float((a[lo:hi] - b[lo:hi]).norm() / a[lo:hi].norm()) <= 1e-4



def test_lr_schedule_survives_host_run_ahead():
    """The per-step learning rate crosses to the device through a ring of pinned slots (with a single
    slot the host, which runs ahead of the GPU in the lagged mode, could rewrite it before its queued copy ran).
    Twelve lagged replays with a steep warm-up schedule; after each call the device-resident lr is copied, in stream
    order, into a log: every step must have used its own lr, and the step counter must end at twelve."""
    from shg_vqa_b200.trainer import HotPathTrainer, warmup_linear
    dev = torch.device("cuda:0")
    h_inp, h_batch = _make(dev, 4)
    inp = {k: v.to(dev) for k, v in h_inp.items()}
    batch = {k: v.to(dev) for k, v in h_batch.items()}
    tr = HotPathTrainer(dev, lr=1e-3, t_total=24, warmup=0.5, dropout=False)
    tr.init_weights(seed=11)
    tr.capture(inp, batch, warmup=1)
    n = 12                                            # more steps than ring slots (8): slots are reused
    log = torch.zeros(n, device=dev)
    for i in range(n):
        tr.train_step_host(h_inp, h_batch, next_inp=h_inp, next_batch=h_batch, lagged=True)
        log[i:i + 1].copy_(tr.lr_dev, non_blocking=True)        # ordered after step i's replay, before step i+1's copy
    tr.drain_losses()
    torch.cuda.synchronize()
    want = torch.tensor([1e-3 * warmup_linear(i / 24, 0.5) for i in range(n)])
    assert tr.step_count == n
    assert torch.allclose(log.cpu(), want, rtol=1e-6, atol=0), (log.cpu(), want)
