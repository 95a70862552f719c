"""Host-side logic of bench.py (no GPU): kernel-name -> entry-point families, the parser that takes `roofline.traffic`
from the committed ncu export, the synthetic batch (shapes the headline config names, CSR targets), and the command
line.  The measured legs themselves run on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_kernel_names_map_to_entry_points():
    assert bench.family_of("void shg::gemm2_kernel<false, false, false, true>(CUtensorMap_st, ...)") == "shg_gemm"
    assert bench.family_of("void shg::gemm_kernel<true, true, true>(...)") == "shg_gemm"
    assert bench.family_of("void shg::attn_fwd2_kernel<false, true>(...)") == "shg_attn_fwd"
    assert bench.family_of("void shg::attn_bwd2_dkv_kernel<2>(...)") == "shg_attn_bwd"
    assert bench.family_of("void shg::attn_bwd_dq_kernel<2, true>(...)") == "shg_attn_bwd"
    assert bench.family_of("void shg::layernorm_bwd_kernel<3>(...)") == "shg_layernorm_bwd_grouped"
    assert bench.family_of("shg::bert_adam_kernel(...)") == "shg_bert_adam_step"
    assert bench.family_of("ncclDevKernel_AllReduce_Sum_f32_RING_LL(...)") == "nccl"
    assert bench.family_of("void at::native::vectorized_elementwise_kernel<...>") == "torch (memset / copies)"


def test_roofline_traffic_comes_from_the_committed_ncu_export():
    got = bench.ncu_traffic("prof_gemm_plain_r*.raw.csv")
    assert got is not None
    nbytes, path, kernel, dur_us = got
    assert "gemm" in kernel and path.endswith(".raw.csv") and os.path.exists(os.path.join(ROOT, path) if not os.path.isabs(path) else path)
    # 12576 x 3072 x 768 bf16: A 19.3 MB + B 4.7 MB read, D 77.3 MB written = 101 MB algorithmic.  The capture may show
    # LESS (operands and part of D stay in the 126 MB L2 between the profiled launches) but never much more: traffic
    # well above the algorithmic bytes would mean wasted re-reads
    algorithmic = 2 * (12576 * 768 + 3072 * 768 + 12576 * 3072)
    assert 0.2 * algorithmic <= nbytes <= 1.25 * algorithmic, (nbytes, algorithmic)
    assert 30.0 < dur_us < 120.0
    assert bench.ncu_traffic("no_such_capture_*.csv") is None


def test_synthetic_batch_has_the_headline_shapes():
    inp, batch = bench.make_batch(1234, pin=False)
    B = bench.BATCH_PER_GPU
    assert B == 32
    assert tuple(inp["visn_tokens"].shape) == (B * 393, 768) and str(inp["visn_tokens"].dtype) == "torch.bfloat16"
    assert inp["input_ids"].shape[0] == B and inp["input_mask"].shape == inp["input_ids"].shape
    assert tuple(batch["target"].shape) == (B, 171)
    for fam in ("rel", "act"):
        off, lab = batch[fam + "_off"], batch[fam + "_labels"]
        assert off.numel() == 16 * B + 1 and int(off[0]) == 0 and int(off[-1]) == lab.numel()
        assert bool((off[1:] >= off[:-1]).all())
    # different seeds (ranks) draw different shards of the same shape
    inp2, _ = bench.make_batch(1235, pin=False)
    assert inp2["visn_tokens"].shape == inp["visn_tokens"].shape
    assert not bool((inp2["input_ids"] == inp["input_ids"]).all())


def test_command_line_and_missing_gpu_fail_loudly():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
    import torch
    if not torch.cuda.is_available():
        # our arm has no CPU fallback: it must exit non-zero with a message, not print a number
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-extras",
                              "--no-cpu-baseline"], capture_output=True, text=True, timeout=300)
        assert out.returncode != 0
        assert not any(l.startswith("{") and '"value"' in l for l in out.stdout.splitlines())
