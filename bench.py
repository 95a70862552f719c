#!/usr/bin/env python
"""Headline benchmark: train clips/s of the SHG-VQA hot path at bs32/GPU on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our path (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                          # the reference's CPU path (oracle port), host cores

A "step" is one full pass of the hot path over one synthetic AGQA-shaped batch (32 clips per GPU): embeddings,
5 language + 5 relation layers, the 2 shared LXRT cross layers + pooler + answer head (forward only, as in the
reference's --taskHGQA training), both 5-layer hypergraph decoders, heads, the hypergraph x question encoder,
per-frame Hungarian matching, set loss + answer loss, the whole backward, the NCCL gradient all-reduce (N>1),
grad-norm clipping and BertAdam.  The slow_r50 backbone stays on stock PyTorch and is OUTSIDE the measured path
(SURVEY.md §8a V*, §2 #9): the headline step starts from the [B,393,768] visual tokens; the same step started from
the backbone's features (through the Conv3d front-end kernels) and from raw clips (through a stock slow_r50) are
reported beside it as `with_frontend` / `with_backbone`.

Where the per-kernel numbers come from: CUPTI kernel records (torch.profiler) of ONE REPLAY of the captured step —
pure device durations inside the real step, no host enqueue gaps — summed per kernel family; the per-shape table of
the dominant family comes from CUPTI records of one eager step joined to the launch order through correlation
ids.  `roofline.traffic` is parsed from the committed `ncu --set full` capture under profiles/ (never a constant).
"""
import argparse
import csv
import glob
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH_PER_GPU = 32
TRAIN_GFLOP_PER_CLIP = 199.3      # 3 x 62.29 (fwd+bwd blocks) + 12.45 (forward-only LXRT x-layers)
FWD_GFLOP_PER_CLIP = 74.74
METRIC = "train clips/sec at bs32/GPU (hot path)"
WORKLOAD = ("AGQA taskHGQA hot-path train step, bs32/GPU, from [B,393,768] visual tokens + question ids: "
            "l5/r5/x2 LXRT stack, 2x5-layer hypergraph decoders, heads, CrossEncoder, per-frame Hungarian matching, "
            "set+answer loss, backward, grad clip, BertAdam (slow_r50 excluded: stock PyTorch; the same step started from "
            "backbone features through the Conv3d front-end is reported as with_frontend)")


def base_config(world):
    """The workload description both arms (ours / --impl reference) print."""
    return {"workload": WORKLOAD, "global_batch": BATCH_PER_GPU * world, "per_gpu_batch": BATCH_PER_GPU,
            "parallelism": "dp%d" % world}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor=p["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed (-lms 100) during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def summary(self):
        samples = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            for line in out.strip().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 7:
                    samples.append(f)
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(s[0])) for s in samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(samples[0][1])), "reasons": reasons,
                "power_w_max": max(float(s[2]) for s in samples), "samples": len(sm)}


def make_batch(seed, pin, batch=BATCH_PER_GPU, **kw):
    import torch
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import csr_from_padded
    raw = make_inputs(batch, seed=seed, with_feat=False, visn_tokens=True, **kw)
    inp = {k: raw[k] for k in ("input_ids", "input_mask", "segment_ids", "rel_segment_ids", "act_segment_ids")}
    inp["visn_tokens"] = raw["visn_tokens"].reshape(-1, 768).to(torch.bfloat16)
    b = {"target": raw["target"]}
    b["rel_labels"], b["rel_off"] = csr_from_padded(raw["rel_tokens"], raw["rel_lengths"])
    b["act_labels"], b["act_off"] = csr_from_padded(raw["act_tokens"], raw["act_lengths"])
    if pin:
        inp = {k: v.contiguous().pin_memory() for k, v in inp.items()}
        b = {k: v.contiguous().pin_memory() for k, v in b.items()}
    return inp, b


# ----------------------------------------------------------------------------------------- CUPTI kernel records
# kernel-name fragment -> the C-ABI entry point that launches it (the key of the `kernels` table)
FAMILIES = [("gemm_kernel", "shg_gemm"), ("gemm2_kernel", "shg_gemm"), ("attn_fwd", "shg_attn_fwd"),
            ("attn_bwd", "shg_attn_bwd"), ("layernorm_fwd", "shg_layernorm_fwd_grouped"),
            ("layernorm_bwd", "shg_layernorm_bwd_grouped"), ("colsum", "shg_colsum_accum"),
            ("bert_adam", "shg_bert_adam_step"), ("grad_sumsq", "shg_grad_sumsq"), ("copy3d", "shg_copy3d"),
            ("embed_scatter", "shg_embed_scatter_bwd"), ("embed_ln", "shg_embed_ln_fwd"), ("f32_to_bf16", "shg_f32_to_bf16"),
            ("act_bwd", "shg_act_bwd"), ("match_from_logits", "shg_match_from_logits"), ("set_loss_fwd", "shg_set_loss_fwd"),
            ("set_loss_bwd", "shg_set_loss_bwd"), ("bce_logits", "shg_bce_logits"), ("conv_gemm", "shg_conv3d"),
            ("conv_wgrad", "shg_conv3d"), ("ncs_to_nsc", "shg_layout"), ("nsc_to_ncs", "shg_layout"),
            ("frontend_", "shg_frontend"), ("add_rows", "shg_add_rows"), ("nccl", "nccl")]


def family_of(kernel_name):
    for frag, fam in FAMILIES:
        if frag in kernel_name:
            return fam
    return "torch (memset / copies)"


def cupti_records(fn, pre=None):
    """Runs fn() under torch.profiler; returns (kernel records sorted by start, launch-ordered correlation ids).
    A kernel record is (start_us, dur_us, name, correlation id).  pre() runs inside the profiler just before fn():
    data-parallel ranks meet at a barrier there, so that no rank's NCCL kernels spin waiting for a peer that is
    still starting its profiler."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        if pre is not None:
            pre()
        fn()
        torch.cuda.synchronize()
    kernels, launches = [], []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            if e.time_range.end > e.time_range.start and "memcpy" not in e.name.lower() and "memset" not in e.name.lower():
                kernels.append((e.time_range.start, e.time_range.end - e.time_range.start, e.name, e.id))
        elif e.name.startswith(("cudaLaunchKernel", "cuLaunchKernel")):
            launches.append((e.time_range.start, e.id))
    kernels.sort()
    launches.sort()
    return kernels, [cid for _, cid in launches]


def kernel_table_from_replay(tr, work_by_entry, pre=None):
    """CUPTI durations of one graph replay summed per entry point + the step's concurrency profile."""
    import torch
    for _ in range(2):
        tr.train_step_graph()
    torch.cuda.synchronize()

    # three replays back to back, the LAST one analysed: the first absorbs the ranks' different profiler start-up
    # times (its all-reduces wait for the slowest rank), after it the ranks run in step
    reps = 3
    kernels, _ = cupti_records(lambda: [tr.train_step_graph() for _ in range(reps)], pre)
    # the barrier the ranks met at is itself a small NCCL kernel: it ends before the step's first compute kernel starts
    first_compute = min(s for s, d, n, _ in kernels if "nccl" not in n)
    kernels = [k for k in kernels if not ("nccl" in k[2] and k[0] + k[1] <= first_compute)]
    if len(kernels) % reps == 0:          # replays of one graph on one stream do not overlap: equal thirds by start time
        kernels = kernels[-(len(kernels) // reps):]
    agg = {}
    for s, d, name, _ in kernels:
        a = agg.setdefault(family_of(name), dict(launches=0, us=0.0))
        a["launches"] += 1
        a["us"] += d
    t0, t1 = kernels[0][0], max(s + d for s, d, _, _ in kernels)
    pts = sorted([(s, 1) for s, d, _, _ in kernels] + [(s + d, -1) for s, d, _, _ in kernels])
    conc, last, hist = 0, t0, {}
    for t, dd in pts:
        if t > last:
            hist[conc] = hist.get(conc, 0.0) + (t - last)
        conc += dd
        last = t
    total = sum(a["us"] for a in agg.values())
    # NCCL time that no compute kernel covers (the exposed part of the path's one collective)
    ev = sorted([(s, 0, "nccl" in n) for s, d, n, _ in kernels] + [(s + d, 1, "nccl" in n) for s, d, n, _ in kernels])
    n_nccl = n_comp = 0
    last, nccl_us, exposed_us = t0, 0.0, 0.0
    for t, end, is_nccl in ev:
        if t > last:
            if n_nccl:
                nccl_us += t - last
                if not n_comp:
                    exposed_us += t - last
        last = t
        if is_nccl:
            n_nccl += -1 if end else 1
        else:
            n_comp += -1 if end else 1
    table = {}
    for fam, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        w = work_by_entry.get(fam)
        row = dict(launches=a["launches"], ms=round(a["us"] * 1e-3, 3), share=round(a["us"] / total, 3))
        if w is not None and w["work"] > 0:
            unit = 1e12 if w["kind"] == "tensor" else 1e9
            row["rate"] = round(w["work"] / (a["us"] * 1e-6) / unit, 1)
            row["unit"] = "TFLOP/s" if w["kind"] == "tensor" else "GB/s"
            row["bound"] = w["kind"]
        table[fam] = row
    prof = {"kernels_in_replay": len(kernels), "span_ms": round((t1 - t0) * 1e-3, 3),
            "sum_of_kernel_durations_ms": round(total * 1e-3, 3),
            "ms_at_concurrency": {str(k): round(v * 1e-3, 3) for k, v in sorted(hist.items())}}
    if nccl_us:
        prof["nccl_active_ms"] = round(nccl_us * 1e-3, 3)
        prof["nccl_exposed_ms"] = round(exposed_us * 1e-3, 3)
    return table, prof


def shape_table_from_eager(tr, inp, batch, entry):
    """Per-shape times of `entry` (e.g. shg_gemm): CUPTI kernel durations of one EAGER step, attributed to the
    calls through the launch order (correlation ids).  Returns [] if the join does not line up."""
    from shg_vqa_b200 import _lib
    prof = _lib.Profiler()
    _lib.profiler = prof
    try:
        kernels, order = cupti_records(lambda: tr.train_step(inp, batch))
    finally:
        _lib.profiler = None
    by_cid = {cid: (d, name) for _, d, name, cid in kernels if "shg::" in name}
    ordered = [by_cid[c] for c in order if c in by_cid]
    need = sum(r[5] for r in prof.calls)
    if need != len(ordered):
        return [], "launch-order join failed: %d kernels expected, %d recorded" % (need, len(ordered))
    agg, i = {}, 0
    for name, kind, work, tag, _, n_k in prof.calls:
        d = sum(x[0] for x in ordered[i:i + n_k])
        i += n_k
        if name == entry:
            a = agg.setdefault(tag, dict(calls=0, us=0.0, work=0.0, kind=kind))
            a["calls"] += 1
            a["us"] += d
            a["work"] += work
    return agg, None


def ncu_traffic(pattern):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest committed `ncu --page raw --csv`
    export matching profiles/<pattern>; (bytes, file, kernel name, duration_us) or None."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=os.path.getmtime)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in reversed(files):
        try:
            rows = list(csv.reader(open(path)))
            hdr, units, data = rows[0], rows[1], rows[2:]
            ir, iw, ik, it = (hdr.index(c) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum", "Kernel Name",
                                                      "gpu__time_duration.sum"))
            tot = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in data]
            dur = [float(r[it]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[it], 1.0) for r in data]
            return sum(tot) / len(tot), os.path.relpath(path, ROOT), data[0][ik].split("(")[0], sum(dur) / len(dur)
        except Exception:
            continue
    return None


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from shg_vqa_b200 import _lib
    from shg_vqa_b200.trainer import HotPathTrainer

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # train mode as the reference runs it (model.train()): dropout live at every site of the path
    tr = HotPathTrainer(dev, world_size=world, t_total=100000, warmup=0.1, lr=1e-5, dropout=not args.no_dropout,
                        dropout_seed=20261018, rank=rank)
    tr.init_weights(seed=9595)                      # identical replicas: same seed on every rank
    h_inp, h_batch = make_batch(1234 + rank, pin=True)
    inp = {k: v.to(dev) for k, v in h_inp.items()}
    batch = {k: v.to(dev) for k, v in h_batch.items()}
    h2d = sum(v.numel() * v.element_size() for v in list(h_inp.values()) + list(h_batch.values()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---------------------------------------------------------------- device-resident throughput (`value`)
    # the step (~650 launches over several streams, the NCCL all-reduces included) is captured once and replayed
    step_fn = (lambda: tr.train_step(inp, batch))
    if not args.eager:
        tr.capture(inp, batch)
        step_fn = tr.train_step_graph
    for _ in range(args.warmup):
        step_fn()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        out, scal = step_fn()
    e.record()
    barrier()
    clocks = sampler.summary() if rank == 0 else None       # sampled over the timed region only
    launches = _lib.launch_count - l0
    if not args.eager:
        launches = tr.launches_per_step * args.steps        # kernels inside the replayed graph
    ms_step = max_over_ranks(s.elapsed_time(e) / args.steps)
    losses = tr.losses_from_scalars(scal)

    # ---------------------------------------------------------------- end to end from pinned host buffers (`e2e`)
    e2e_ms = None
    if not args.eager:
        for _ in range(2):                          # untimed: staging buffers, copy stream, pinned loss slots
            tr.train_step_host(h_inp, h_batch, next_inp=h_inp, next_batch=h_batch, lagged=True)
        tr.drain_losses()
        barrier()
        t0 = time.perf_counter()
        s.record()
        n_read = 0
        for _ in range(args.steps):
            # H2D of this step's inputs + step + D2H of the loss sums, every step.  The H2D of the FOLLOWING step's
            # inputs (the same pinned buffers here, a data loader's next batch in use) is started on a copy stream
            # while this step computes, and each step's loss is read back asynchronously and handed out one call
            # later (as a training loop that logs a lagged loss does): every step's inputs cross PCIe once and every
            # step's result reaches the host inside the timed region.
            n_read += tr.train_step_host(h_inp, h_batch, next_inp=h_inp, next_batch=h_batch, lagged=True) is not None
        n_read += tr.drain_losses() is not None
        assert n_read == args.steps, n_read
        e.record()
        barrier()
        e2e_ms = max_over_ranks(max(s.elapsed_time(e), (time.perf_counter() - t0) * 1e3) / args.steps)

    # ---------------------------------------------------------------- data-parallel correctness (outside the timed region)
    dp = dp_check(tr, dev, rank, world) if world > 1 else None

    # ---------------------------------------------------------------- per-kernel numbers
    # every rank runs the same replays / eager steps (the step contains collectives); rank 0 reports
    peaks = read_peaks()
    prof = _lib.Profiler()                          # algorithmic work per entry point: one eager step (times unused)
    _lib.profiler = prof
    tr.train_step(inp, batch)
    _lib.profiler = None
    torch.cuda.synchronize()
    work = prof.work_by_entry()
    kernels, step_profile, shapes_note = {}, None, None
    try:
        kernels, step_profile = kernel_table_from_replay(tr, work, barrier if world > 1 else None) if not args.eager else ({}, None)
    except Exception as ex:                         # CUPTI unavailable: say so instead of inventing numbers
        shapes_note = "CUPTI kernel records unavailable (%s: %s)" % (type(ex).__name__, str(ex)[:120])
        if world > 1:
            raise
    ranked = [(k, v) for k, v in kernels.items() if "rate" in v]
    roof = {"bound": None, "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None}
    top = max(ranked, key=lambda kv: kv[1]["ms"])[0] if ranked else "shg_gemm"
    agg, err = shape_table_from_eager(tr, inp, batch, top)
    line = None
    if rank == 0:
        if ranked:
            name, a = top, kernels[top]
            peak = peaks["tensor"] if a["bound"] == "tensor" else peaks["hbm"]
            roof = dict(bound="tensor" if a["bound"] == "tensor" else "hbm", kernel=name, achieved=a["rate"], peak=peak,
                        unit=a["unit"], frac=round(a["rate"] / peak, 3),
                        peak_source=peaks["src"] + (" (sustained: kernel timed inside a long step)" if a["bound"] == "tensor" else ""),
                        timing="CUPTI kernel durations of one graph replay, summed over the %d launches" % a["launches"],
                        launches_per_step=a["launches"], share_of_kernel_time=a["share"],
                        avg_launch_us=round(1e3 * a["ms"] / a["launches"], 2), traffic=None)
            tfile = {"shg_gemm": "prof_gemm_plain_r*.raw.csv", "shg_attn_fwd": "prof_attn_fwd_r*.raw.csv",
                     "shg_attn_bwd": "prof_attn_bwd_r*.raw.csv"}.get(name)
            t = ncu_traffic(tfile) if tfile else None
            if t:
                roof["traffic"] = round(t[0])
                roof["traffic_of"] = "%s, one launch of %s (%.1f us under ncu), dram__bytes_read.sum + dram__bytes_write.sum" % (
                    t[1], t[2], t[3])
            if err:
                shapes_note = err
            shapes = []
            for tag, x in sorted(agg.items(), key=lambda kv: -kv[1]["us"])[:10] if agg else []:
                unit = 1e12 if x["kind"] == "tensor" else 1e9
                r = x["work"] / (x["us"] * 1e-6) / unit
                shapes.append(dict(shape=tag, calls=x["calls"], avg_launch_us=round(x["us"] / x["calls"], 2),
                                   achieved=round(r, 1), frac=round(r / peak, 3)))
            roof["largest_shapes"] = shapes
        if shapes_note:
            roof["note"] = shapes_note
        clips = BATCH_PER_GPU * world
        cfg = base_config(world)
        cfg.update({"dropout": "off" if args.no_dropout else "train mode: hidden/attention-prob 0.1, decoders 0.15, fused counter-based masks",
                    "launch": "eager" if args.eager else "cuda-graph replay",
                    "l2": "no flush needed: each step streams ~5 GB of parameters/moments/activations (>> 126 MB L2)",
                    "hot_path_tflop_per_step_per_gpu": round(TRAIN_GFLOP_PER_CLIP * BATCH_PER_GPU / 1e3, 2)})
        if world > 1:
            cfg["collective"] = ("fp32 NCCL all-reduce of the 191 M-element live gradient span in 6 groups, each issued when its "
                                 "gradients are final (late+language+embeddings | r4 | r3 | r2 | r1 | r0+vectors), captured in the step's graph")
        line = {
            "metric": METRIC, "value": round(clips / (ms_step * 1e-3), 2),
            "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "model_flops_utilisation": round(TRAIN_GFLOP_PER_CLIP * BATCH_PER_GPU / 1e3 / (ms_step * 1e-3) / peaks["tensor"], 3),
            "kernels": kernels,
            "step_profile": step_profile,
            "loss": {k: round(v, 4) for k, v in losses.items()},
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": round(clips / (e2e_ms * 1e-3), 2), "unit": "clips/s", "h2d_bytes_per_step": int(h2d),
                           "d2h_bytes_per_step": 64, "ms_per_step": round(e2e_ms, 3)}
        if dp is not None:
            line["dp_check"] = dp
    if world > 1:
        dist.barrier()
    if rank != 0:
        shutdown(tr, world)
        return
    if world == 1 and not args.eager:
        line["fwd_ms_per_batch"] = measure_forward(tr, inp, batch, args)
    if world == 1:
        tr.release()
        del tr
        torch.cuda.empty_cache()
    if world == 1 and not args.no_extras:
        extras = {}
        for name, fn in (("with_frontend", lambda: measure_wider(args, dev, peaks, backbone=False)),
                         ("with_backbone", lambda: measure_wider(args, dev, peaks, backbone=True)),
                         ("config4_vqa_eval_bs8", lambda: measure_config4(args, dev)),
                         ("config5_star_bs64", lambda: measure_config5(args, dev)),
                         ("gpu_eager_reference", lambda: gpu_eager_reference(dev, ms_step))):
            try:
                extras[name] = fn()
            except Exception as ex:
                extras[name] = {"error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
            torch.cuda.empty_cache()
        line["with_frontend"] = extras.pop("with_frontend")
        line["extra"] = extras
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_throughput(steps=2, warmup=1)
        if not args.no_extras:
            line["cpu_baseline"]["other_configs"] = cpu_other_configs()
    print(json.dumps(line), flush=True)
    if world > 1:
        shutdown(tr, world)


def shutdown(tr, world):
    """Orderly end of a data-parallel rank: the captured graph holds NCCL resources and must go before the process
    group does; a timer guards against a communicator teardown that never returns (the JSON line is out by then)."""
    import threading
    import torch.distributed as dist
    if world <= 1:
        return
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    tr.release()
    dist.destroy_process_group()
    t.cancel()


def measure_forward(tr, inp, batch, args):
    """BASELINE.json's second metric, `fwd ms/batch`: forward + matching + losses of the bs32 train batch (train
    mode: dropout on, activations saved), captured and replayed; CUDA events over args.steps replays."""
    import torch
    eng = tr.engine
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            out, ctx = eng.forward(inp, train=True)
            tr.loss_and_grads(out, batch, want_grad=False)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out, ctx = eng.forward(inp, train=True)
        tr.loss_and_grads(out, batch, want_grad=False)
    for _ in range(3):
        g.replay()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(args.steps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    peaks = read_peaks()
    del g
    return {"value": round(ms, 3), "unit": "ms", "what": "forward + device matching + losses, bs32, train mode, graph replay",
            "tflops": round(FWD_GFLOP_PER_CLIP * BATCH_PER_GPU / 1e3 / (ms * 1e-3), 1),
            "frac_of_sustained_tensor_peak": round(FWD_GFLOP_PER_CLIP * BATCH_PER_GPU / 1e3 / (ms * 1e-3) / peaks["tensor"], 3)}


def dp_check(tr, dev, rank, world):
    """Data-parallel numerics on the hardware, outside the timed region:
    (1) after the captured steps every rank holds bit-identical weights;
    (2) dropout off, global CE normalisation on (the reference's nn.DataParallel normalises the weighted CE over
        the gathered batch, tasks/agqaHGQA.py:127,223): the all-reduced gradient of `world` shards of 4 clips / world
        equals the single-process gradient of the concatenated batch."""
    import torch
    import torch.distributed as dist
    from shg_vqa_b200.trainer import HotPathTrainer
    live = tr.store.live_extent
    w = tr.store.master[live]
    digest = torch.stack([w.double().sum(), w.double().abs().sum(),
                          w.view(torch.int32).to(torch.int64).sum().double()])
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(torch.equal(all_d[0], d) for d in all_d)
    # gradient equivalence at a small per-rank batch (eager steps, no optimiser)
    per = 4
    chk = HotPathTrainer(dev, world_size=world, dropout=False, global_ce_norm=True, rank=rank)
    chk.init_weights(seed=77)
    shards = [make_batch(4321 + r, pin=False, batch=per) for r in range(world)]
    inp = {k: v.to(dev) for k, v in shards[rank][0].items()}
    bat = {k: v.to(dev) for k, v in shards[rank][1].items()}
    out, ctx = chk.engine.forward(inp, train=True)
    _, d_hg, d_rel, d_act, _ = chk.loss_and_grads(out, bat)
    ce_sums = chk.scal_den.clone()                 # the globally reduced CE sums every rank normalised with
    chk.engine.backward(ctx, d_hg, d_rel, d_act)
    chk.all_reduce_grads()
    torch.cuda.synchronize()
    g_dp = chk.store.grad[chk.store.live_extent] / world
    rel = None
    if rank == 0:
        one = HotPathTrainer(dev, world_size=1, dropout=False)
        one.init_weights(seed=77)
        cat_inp = {k: torch.cat([s[0][k] for s in shards]).to(dev) for k in shards[0][0]}
        cat_bat = {"target": torch.cat([s[1]["target"] for s in shards]).to(dev)}
        for fam in ("rel", "act"):
            labs = [s[1][fam + "_labels"] for s in shards]
            offs, base = [shards[0][1][fam + "_off"][:1]], 0
            for s_ in shards:
                o = s_[1][fam + "_off"]
                offs.append(o[1:] + base)
                base += int(o[-1])
            cat_bat[fam + "_labels"] = torch.cat(labs).to(dev)
            cat_bat[fam + "_off"] = torch.cat(offs).to(torch.int32).to(dev)
        out, ctx = one.engine.forward(cat_inp, train=True)
        # same normaliser bits as the ranks used: the concatenated batch sums the same weights in another order, and
        # a last-bit difference in that denominator alone moves the gradient by ~4e-4 rel-L2 through the bf16
        # roundings of the backward (tools/uninit_probe.py, tools/race_probe.py: measured on the single-process
        # step before the loss sums were made order-independent)
        _, d_hg, d_rel, d_act, _ = one.loss_and_grads(out, cat_bat, ce_sums=ce_sums)
        one.engine.backward(ctx, d_hg, d_rel, d_act)
        torch.cuda.synchronize()
        g_one = one.store.grad[one.store.live_extent]
        rel = float((g_dp.double() - g_one.double()).norm() / g_one.double().norm())
        del one
    del chk
    torch.cuda.empty_cache()
    res = {"weights_identical_across_ranks": bool(same), "ranks": world,
           "reduced_grad_vs_single_process_rel_l2": rel,
           "how": "%d clips per rank, dropout off, global CE normalisation, the single-process run normalised with "
                  "the ranks' reduced CE sums; tolerance 1e-3 = the bf16 noise floor of the backward (one flipped "
                  "bf16 rounding anywhere moves the gradient by 1e-4-level rel-L2); fp32 split-K sums run in a "
                  "different order on a %dx larger M" % (per, world)}
    if rank == 0:
        # reported, not raised: a failed check must not cost the run its bench line (tests/test_gpu_dp.py asserts it)
        res["ok"] = bool(same and rel is not None and rel <= 1e-3)
        if not res["ok"]:
            print("bench.py: DATA-PARALLEL CHECK FAILED: weights identical %s, gradient rel-L2 %r" % (same, rel),
                  file=sys.stderr, flush=True)
    return res


def _time_graph_step(tr, inp, batch, args):
    import torch
    tr.capture(inp, batch)
    for _ in range(args.warmup):
        tr.train_step_graph()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        out, scal = tr.train_step_graph()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / args.steps, scal


def measure_wider(args, dev, peaks, backbone):
    """The same train step started earlier.  backbone=False: from the backbone's fp32
    [B,2048,16,7,7] features through the Conv3d front-end as tcgen05 implicit GEMM (forward, dgrad, both wgrads; its
    97 M parameters join the optimiser's live span).  backbone=True (§8f-4, config 2 as written): from raw clips
    [B,3,16,224,224] through a stock-PyTorch slow_r50 (local definition, bf16 autocast + channels_last_3d, frozen in
    eval mode as STAR runs it, star.py:87-88) whose output feeds the same step."""
    import torch
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import HotPathTrainer, csr_from_padded
    tr = HotPathTrainer(dev, t_total=100000, warmup=0.1, lr=1e-5, dropout=not args.no_dropout,
                        dropout_seed=20261018, frontend=True)
    tr.init_weights(seed=9595)
    raw = make_inputs(BATCH_PER_GPU, seed=1234, with_feat=True)
    inp = {k: raw[k].to(dev) for k in ("feat", "input_ids", "input_mask", "segment_ids", "rel_segment_ids",
                                       "act_segment_ids")}
    batch = {"target": raw["target"].to(dev)}
    batch["rel_labels"], batch["rel_off"] = [t.to(dev) for t in csr_from_padded(raw["rel_tokens"], raw["rel_lengths"])]
    batch["act_labels"], batch["act_off"] = [t.to(dev) for t in csr_from_padded(raw["act_tokens"], raw["act_lengths"])]
    ms, scal = _time_graph_step(tr, inp, batch, args)
    losses = tr.losses_from_scalars(scal)
    res = {"value": round(BATCH_PER_GPU / (ms * 1e-3), 2), "unit": "clips/s", "ms_per_step": round(ms, 3),
           "loss_total": round(losses["total"], 4)}
    if not backbone:
        res["workload"] = ("same step from backbone features fp32 [32,2048,16,7,7]: + layout cast, Conv3d(2048->768,(5,3,3))"
                           "+GELU, Conv3d(768->768)+GELU, cls/position/dropout, their backward (no gradient into the "
                           "frozen-or-stock backbone), +97 M live parameters")
        return res
    from shg_vqa_b200.backbone import VideoBackbone
    g = torch.Generator().manual_seed(5)
    clip = torch.randn(BATCH_PER_GPU, 3, 16, 224, 224, generator=g).to(dev)
    net = VideoBackbone("slow_r50", autocast=True, freeze=True).to(dev)

    def encode():
        return net.encode(clip)
    for _ in range(2):
        f = encode()
    assert tuple(f.shape) == (BATCH_PER_GPU, 2048, 16, 7, 7), tuple(f.shape)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        f = encode()
        tr.static_inp["feat"].copy_(f)
        tr.train_step_graph()
    e.record()
    torch.cuda.synchronize()
    ms_all = s.elapsed_time(e) / args.steps
    res.update({"workload": "config 2 as written: raw clips fp32 [32,3,16,224,224] -> stock-PyTorch slow_r50 (bf16 autocast, "
                            "channels_last_3d, frozen/eval as star.py:87-88, random init) -> features -> the with_frontend step",
                "value": round(BATCH_PER_GPU / (ms_all * 1e-3), 2), "ms_per_step": round(ms_all, 3),
                "backbone_ms": round(ms_all - ms, 3), "hot_path_and_frontend_ms": round(ms, 3)})
    return res


def measure_config4(args, dev):
    """Config 4: --taskVQA eval bs8 (tasks/agqa_model.py:185-192): features -> Conv3d front-end -> language /
    relation / cross layers -> pooler -> answer head; no decoders, no matcher; no_grad; graph replay."""
    import torch
    from shg_vqa_b200.synthetic import make_inputs
    from shg_vqa_b200.trainer import HotPathTrainer
    tr = HotPathTrainer(dev, task="vqa", frontend=True)
    tr.init_weights(seed=9595)
    raw = make_inputs(8, seed=4, with_feat=True)
    inp = {k: raw[k].to(dev) for k in ("feat", "input_ids", "input_mask", "segment_ids")}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            tr.eval_step(inp)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = tr.eval_step(inp)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / args.steps
    return {"workload": "AGQA --taskVQA eval, bs8, from backbone features (Conv3d front-end + l5/r5/x2 + pooler + answer head), "
                        "no_grad, graph replay", "value": round(8 / (ms * 1e-3), 1), "unit": "clips/s",
            "fwd_ms_per_batch": round(ms, 3), "logit_shape": list(out["logit"].shape)}


def measure_config5(args, dev):
    """Config 5: STAR-shaped decoder + matcher stress, bs64, every frame full (1024 8x8 + 1024 3x3 LSAPs), class
    counts 564 / 112 / 4 (star_model.py:19, star.py:90): the full train step, and the two matcher launches alone."""
    import torch
    from shg_vqa_b200 import ops
    from shg_vqa_b200.trainer import HotPathTrainer
    tr = HotPathTrainer(dev, n_answers=4, n_rel_cls=564, n_act_cls=112, t_total=100000, dropout=not args.no_dropout)
    tr.init_weights(seed=9595)
    h_inp, h_batch = make_batch(77, pin=False, batch=64, n_rel_cls=564, n_act_cls=112, n_ans=4, full_frames=True)
    inp = {k: v.to(dev) for k, v in h_inp.items()}
    batch = {k: v.to(dev) for k, v in h_batch.items()}
    ms, scal = _time_graph_step(tr, inp, batch, args)
    out = tr.static_out
    # matcher alone (softmax stats + cost gather + 1024 LSAPs + target scatter), both families
    for _ in range(3):
        ops.match_from_logits(out["rel_logits"], batch["rel_labels"], batch["rel_off"], 8)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        ops.match_from_logits(out["rel_logits"], batch["rel_labels"], batch["rel_off"], 8)
        ops.match_from_logits(out["act_logits"], batch["act_labels"], batch["act_off"], 3)
    e.record()
    torch.cuda.synchronize()
    m_us = s.elapsed_time(e) / 20 * 1e3
    n_bytes = out["rel_logits"].numel() * 4 + out["act_logits"].numel() * 4
    return {"workload": "STAR-shaped (564/112/4 classes) train step, bs64, all 2048 frames full, graph replay",
            "value": round(64 / (ms * 1e-3), 1), "unit": "clips/s", "ms_per_step": round(ms, 3),
            "matcher_alone_us": round(m_us, 1), "matcher_lsaps": 2048,
            "matcher_logit_bytes": int(n_bytes), "matcher_gbs": round(n_bytes / (m_us * 1e-6) / 1e9, 1),
            "loss_total": round(tr.losses_from_scalars(scal)["total"], 4)}


def gpu_eager_reference(dev, our_ms):
    """SURVEY §8d's honest competitor: the reference's algorithm in stock eager PyTorch ON THE SAME B200 — the oracle
    port (oracle/shg_oracle.py, pinned against the unmodified reference) with tensors on the GPU: forward, matching
    (its CPU scipy-equivalent loop, as matcher.py:76-80 does), losses, autograd backward, clip_grad_norm_ and a
    BertAdam step written with torch foreach ops; fp32 / TF32 / bf16 autocast; bs32, from visual tokens (the same
    workload as the headline).  Baseline leg: the oracle is only timed here, never used by the product path."""
    import torch
    from oracle import shg_oracle as so
    from shg_vqa_b200.params import build_key_shapes
    from shg_vqa_b200.synthetic import make_inputs
    ks = {k: [list(s), d] for k, (s, d) in build_key_shapes("hgqa").items() if ".visn_fc.conv" not in k}
    sd = so.make_state_dict(ks, seed=0)
    leaves = {}
    for k, v in sd.items():
        ck = so.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = v.to(dev).requires_grad_(v.is_floating_point())
    full = {k: leaves[so.canonical_key(k)] for k in sd}
    raw = make_inputs(BATCH_PER_GPU, seed=1234, with_feat=False, visn_tokens=True)
    inp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in raw.items()}
    params = [t for t in leaves.values() if t.is_floating_point()]
    m = [torch.zeros_like(t) for t in params]
    v2 = [torch.zeros_like(t) for t in params]

    def step(autocast):
        for t in params:
            t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = so.agqa_forward(full, inp)
        out = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in out.items()}
        so.hgqa_total_loss(out, inp)["total"].backward()
        live = [(t, a, b) for t, a, b in zip(params, m, v2) if t.grad is not None]
        torch.nn.utils.clip_grad_norm_([t for t, _, _ in live], 5.0)
        with torch.no_grad():
            ts, gs = [t for t, _, _ in live], [t.grad for t, _, _ in live]
            ms_, vs_ = [a for _, a, _ in live], [b for _, _, b in live]
            torch._foreach_mul_(ms_, 0.9)
            torch._foreach_add_(ms_, gs, alpha=0.1)
            torch._foreach_mul_(vs_, 0.999)
            torch._foreach_addcmul_(vs_, gs, gs, value=0.001)
            den = torch._foreach_sqrt(vs_)
            torch._foreach_add_(den, 1e-6)
            upd = torch._foreach_div(ms_, den)
            torch._foreach_add_(upd, ts, alpha=0.01)
            torch._foreach_add_(ts, upd, alpha=-1e-5)
    res = {}
    for name, tf32, ac in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        step(ac)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            step(ac)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        res[name] = {"ms_per_step": round(ms, 2), "clips_per_s": round(BATCH_PER_GPU / (ms * 1e-3), 1),
                     "hot_path_speedup": round(ms / our_ms, 2)}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    res["what"] = ("oracle port of the reference's algorithm in stock eager PyTorch on this GPU: fwd + matching (CPU LSAP loop, "
                   "as the reference) + loss + autograd bwd + clip + BertAdam (foreach), bs32 from visual tokens; wall clock "
                   "over 3 steps after 1 warm-up")
    return res


# -------------------------------------------------------------------------------------- CPU reference arm
def _oracle_leaves(task="hgqa", frontend=False):
    import torch
    from oracle import shg_oracle as so
    from shg_vqa_b200.params import build_key_shapes
    ks = {k: [list(s), d] for k, (s, d) in build_key_shapes(task).items()}
    if not frontend:
        ks = {k: v for k, v in ks.items() if ".visn_fc.conv" not in k}          # front-end is outside the path
    sd = so.make_state_dict(ks, seed=0)
    leaves = {}
    for k, v in sd.items():
        ck = so.canonical_key(k)
        if ck not in leaves:
            leaves[ck] = v.clone().requires_grad_(v.is_floating_point())
    return so, leaves, {k: leaves[so.canonical_key(k)] for k in sd}


def cpu_reference_throughput(steps, warmup, sample_clips=BATCH_PER_GPU, budget_s=240.0):
    """The reference's own algorithm on the host cores: oracle/shg_oracle.py (torch fp32 port, pinned against the
    unmodified reference by tests/golden) — forward, scipy-equivalent matching (oracle/lsap.c), loss, backward,
    clip_grad_norm_(5.) + BertAdam — on full bs32 batches of the headline workload (a smaller sample only if the
    requested steps would not fit the time budget; said in `sample`)."""
    import torch
    from shg_vqa_b200.synthetic import make_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    so, leaves, full = _oracle_leaves()
    state = {}

    def make(n):
        return make_inputs(n, seed=1234, with_feat=False, visn_tokens=True)

    def step(inp):
        for t in leaves.values():
            t.grad = None
        out = so.agqa_forward(full, inp)
        so.hgqa_total_loss(out, inp)["total"].backward()
        # clip_grad_norm_(5.) + BertAdam (tasks/agqaHGQA.py:391-392, lxrt/optimization.py:101-180)
        live = [t for t in leaves.values() if t.grad is not None]
        torch.nn.utils.clip_grad_norm_(live, 5.0)
        with torch.no_grad():
            for t in live:
                m, v = state.setdefault(id(t), (torch.zeros_like(t), torch.zeros_like(t)))
                m.mul_(0.9).add_(t.grad, alpha=0.1)
                v.mul_(0.999).addcmul_(t.grad, t.grad, value=0.001)
                t.add_(m / (v.sqrt() + 1e-6) + 0.01 * t, alpha=-1e-5)

    inp = make(sample_clips)
    t0 = time.perf_counter()
    step(inp)                                           # first (warm-up) step also sizes the run
    t_first = time.perf_counter() - t0
    if t_first * (steps + max(warmup, 1)) > budget_s and sample_clips > 8:
        sample_clips = 8
        inp = make(sample_clips)
        step(inp)
    for _ in range(max(warmup - 1, 0)):
        step(inp)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(inp)
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(sample_clips / dt, 3), "unit": "clips/s", "cores": cores, "kind": "port",
            "sample": "%d steps of %d-clip batches of the headline workload (fwd + matching + loss + bwd + clip + BertAdam, "
                      "fp32, %d torch threads), %.2f s/step" % (steps, sample_clips, cores, dt)}


def cpu_other_configs():
    """SURVEY §8d's other CPU lines: config 1 (B=2 fwd+loss+bwd from backbone features), config 4 (--taskVQA eval
    B=8) and the matcher alone at bs32 / bs64 (the reference's softmax -> gather -> .cpu() -> per-frame LSAP loop,
    lxrt/matcher.py:59-80).  Median of 3 after 1 warm-up."""
    import torch
    from shg_vqa_b200.synthetic import make_inputs
    res = {}

    def med(fn, n=3):
        fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2]
    so, leaves, full = _oracle_leaves(frontend=True)
    raw = make_inputs(2, seed=1234, with_feat=True)

    def c1():
        for t in leaves.values():
            t.grad = None
        so.hgqa_total_loss(so.agqa_forward(full, raw), raw)["total"].backward()
    t = med(c1)
    res["config1_b2_fwd_loss_bwd"] = {"s_per_iter": round(t, 3), "clips_per_s": round(2 / t, 2)}
    raw8 = make_inputs(8, seed=4, with_feat=True)

    def c4():
        with torch.no_grad():
            so.agqa_forward(full, raw8, task="vqa")
    t = med(c4)
    res["config4_vqa_eval_b8"] = {"s_per_iter": round(t, 3), "clips_per_s": round(8 / t, 2)}
    g = torch.Generator().manual_seed(9)
    for b, n_rel, n_act, full_frames in ((32, 457, 158, False), (64, 564, 112, True)):
        r = make_inputs(b, seed=3, n_rel_cls=n_rel, n_act_cls=n_act, full_frames=full_frames, with_feat=False)
        rel = torch.randn(b, 128, n_rel, generator=g)
        act = torch.randn(b, 48, n_act, generator=g)
        t = med(lambda: (so.hungarian_match(rel, r["rel_labels"]), so.hungarian_match(act, r["act_labels"])))
        res["matcher_alone_bs%d" % b] = {"ms_per_iter": round(t * 1e3, 2), "lsaps": 2 * 16 * b}
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference_throughput(steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"],
            "unit": "clips/s", "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(1, args.warmup),
            "ms_per_step": round(BATCH_PER_GPU / base["value"] * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(args.gpus),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip with_frontend / with_backbone / configs 4, 5 / GPU eager reference")
    ap.add_argument("--no-frontend", dest="no_extras", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropout", action="store_true", help="eval-style step (p = 0) for A/B timing")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying the CUDA graph")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
