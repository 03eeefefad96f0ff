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
