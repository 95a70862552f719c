"""Statistical properties of the counter-based dropout mask (host restatement of csrc/shg_common.cuh; the GPU tests
in test_gpu_dropout.py hold the kernels to this restatement element by element)."""
import pytest

from _dropmask import keep_attn, keep_flat


@pytest.mark.parametrize("p", [0.1, 0.15])
def test_mask_is_iid_bernoulli(p):
    """keep rate, pair independence and seed/salt independence of the counter-based mask (host restatement)."""
    k1 = keep_flat(123, 7, 4096, 768, p)
    assert abs(k1.mean() - (1 - p)) < 2e-3
    assert abs((k1[:, 0::2] & k1[:, 1::2]).mean() - (1 - p) ** 2) < 3e-3           # the two lanes of a pair
    assert abs((k1[:, :-1] & k1[:, 1:]).mean() - (1 - p) ** 2) < 3e-3               # neighbours
    assert abs((k1[:-1] & k1[1:]).mean() - (1 - p) ** 2) < 3e-3                     # rows
    for other in (keep_flat(124, 7, 4096, 768, p), keep_flat(123, 8, 4096, 768, p)):
        assert abs((k1 & other).mean() - (1 - p) ** 2) < 3e-3


def test_attention_mask_layout():
    k = keep_attn(1, 2, 2, 12, 40, 393, 0.1)
    assert k.shape == (2, 12, 40, 393) and abs(k.mean() - 0.9) < 3e-3
    # rows are different draws
    assert abs((k[:, :, :-1] & k[:, :, 1:]).mean() - 0.81) < 5e-3
