"""SVA oracle against the committed outputs of the reference modules (tests/golden/sva_*.npz).  Runs anywhere."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import sva_oracle
from oracle.make_golden import sva_inputs
from oracle.synth import make_sva_state_dict

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sva_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[4:-4] for p in GOLDEN])
def test_sva_oracle_matches_reference_golden(path):
    z = np.load(path)
    m = json.loads(str(z["meta"]))
    sd = make_sva_state_dict(m["hidden"], m["tower_dims"], m["window_sides"], m["layers"], m["seed"], m["stress"])
    sizes = [tuple(s) for s in m["image_sizes"]]
    tower = sva_inputs(m["tower_dims"], m["window_sides"], m["query_side"], len(sizes), m["seed"])
    got = sva_oracle.sva_frames(sd, tower, sizes, m["query_side"], m["layers"], num_heads=m["heads"])
    assert got.shape == z["out"].shape
    assert float(np.abs(got.numpy() - z["out"]).max()) <= 1e-4


def test_sva_goldens_exist():
    assert len(GOLDEN) >= 2


def test_kv_fold_is_the_same_function():
    """Host logic of SVAConnector: K and V of a tower are computed as one GEMM on the statistics-only normalised
    input with the two LayerNorm affines folded into the weights; in exact arithmetic that equals
    Linear(LayerNorm(x)) for both projections (vision_sampler.py:192-217)."""
    import torch
    import torch.nn.functional as F
    from tdc_video_b200.sva import SVAConnector
    torch.manual_seed(0)
    m = SVAConnector((96, 64), (2, 2), hidden=128, query_side=4, num_layers=2)
    sd = {k: torch.from_numpy(v) for k, v in make_sva_state_dict(128, (96, 64), (2, 2), 2, 7, 1.0).items()}
    m.load_state_dict(sd, strict=True)
    x = torch.randn(50, 128, dtype=torch.float64)
    xhat = F.layer_norm(x, (128,), None, None, 1e-5)
    for li in range(2):
        for t in range(2):
            w_kv, b_kv, ones, zeros = m._kv_folded(li, t)
            assert w_kv.shape == (256, 128) and w_kv.dtype == torch.bfloat16 and b_kv.dtype == torch.float32
            assert torch.equal(ones, torch.ones(128)) and torch.equal(zeros, torch.zeros(128))
            ca = m.vision_sampler_0.layers[li].cross_attn
            got = F.linear(xhat, w_kv.double(), b_kv.double())
            for j, name in enumerate((f"k_proj_{t}", f"v_proj_{t}")):
                ln, lin = getattr(ca, name)
                ref = F.linear(F.layer_norm(x, (128,), ln.weight.double(), ln.bias.double(), 1e-5), lin.weight.double())
                err = (got[:, 128 * j:128 * (j + 1)] - ref).abs().max() / ref.abs().max()
                assert err < 5e-3, (li, t, name, float(err))   # bf16 rounding of the folded weight only
