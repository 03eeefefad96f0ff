"""The CPU restatement (oracle/) against the committed golden outputs of the reference.

Runs everywhere (no GPU, no /root/reference).  Tolerance: both sides are fp32 CPU torch, so
only summation-order noise is allowed: max |diff| <= 2e-5 on O(1) values."""
import numpy as np
import pytest

from oracle import qformer_oracle as oracle
from tests._golden import GoldenCase, golden_names

SMALL = [n for n in golden_names() if not n.startswith("full_")]
FULL = [n for n in golden_names() if n.startswith("full_")]


@pytest.mark.parametrize("name", SMALL + FULL)
def test_oracle_matches_reference_golden(name):
    c = GoldenCase(name)
    hidden = oracle.qformer_forward(c.sd, c.geom, c.inputs["query_embeds"], c.inputs["enc"], c.inputs["input_ids"],
                                    kv_len=c.kv_len)
    comp = oracle.proj_norm(c.sd, hidden, c.K)
    assert hidden.shape == c.hidden.shape and comp.shape == c.compressed.shape
    assert np.abs(hidden.numpy() - c.hidden).max() <= 2e-5
    assert np.abs(comp.numpy() - c.compressed).max() <= 2e-6
    # compressed tokens are unit-norm (F.normalize, cambrian_arch.py:1664-1667)
    assert np.allclose(np.linalg.norm(comp.numpy(), axis=-1), 1.0, atol=1e-5)


def test_golden_set_covers_edge_cases():
    names = set(golden_names())
    for needed in ("small_notext", "small_text", "small_stress_k16", "small_kvlen", "small_k20_ragged_queries",
                   "freq1_speech_style", "full_literal_l194", "full_qwen_l206_text"):
        assert needed in names


def test_avg_pool_bins_match_reference_formula():
    # bins [floor(i*L/K), ceil((i+1)*L/K)) — overlapping for L=156, K=16
    import torch
    L, K, d = 156, 16, 8
    x = torch.arange(L * d, dtype=torch.float32).reshape(1, L, d)
    got = oracle.avg_pool_queries(x, K)[0]
    for i in range(K):
        s, e = (i * L) // K, -((-(i + 1) * L) // K)
        assert torch.allclose(got[i], x[0, s:e].mean(0))
    assert (0 * L) // K == 0 and -((-1 * L) // K) == 10 and (1 * L) // K == 9
