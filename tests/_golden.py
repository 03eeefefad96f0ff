"""Helpers shared by the parity tests: load a golden case and regenerate its weights/inputs."""
import glob
import json
import os

import numpy as np

from oracle.synth import QFormerGeometry, make_inputs, make_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix="qformer_"):
    return sorted(os.path.basename(p)[len(prefix):-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, f"qformer_{name}.npz"))
        self.name = name
        self.meta = json.loads(str(z["meta"]))
        self.geom = QFormerGeometry(**self.meta["geometry"])
        self.hidden = z["hidden"]
        self.compressed = z["compressed"]
        m = self.meta
        self.rows, self.L, self.K, self.T = m["rows"], m["kv_tokens"], m["num_query"], m["num_text"]
        self.kv_len = m["kv_len"]
        self.sd = make_state_dict(self.geom, m["seed"], stress=m["stress"], with_text=self.T > 0)
        self.inputs = make_inputs(self.geom, m["seed"], self.rows, self.L, self.K, self.T,
                                  audio_tokens=m["audio_tokens"])
