"""Pin the oracle against the UNMODIFIED reference modules, live (build container only).

Skipped where /root/reference is absent (e.g. the GPU box) — there the committed golden
vectors (tests/test_oracle_golden.py), which this same comparison generated, stand in."""
import numpy as np
import pytest
import torch

from oracle import qformer_oracle as oracle
from oracle import ref_shim
from oracle.synth import QFormerGeometry, make_inputs, make_state_dict

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("K,T,L,stress", [(8, 0, 21, 0.0), (16, 7, 33, 8.0), (3, 2, 5, 8.0)])
def test_oracle_equals_reference_bertmodel(K, T, L, stress):
    from oracle.make_golden import run_reference
    geom = QFormerGeometry(hidden=128, heads=2, intermediate=192, layers=3, cross_freq=2, d_enc=72, d_out=96,
                           vocab=50, max_pos=16)
    sd = make_state_dict(geom, 100 + K, stress=stress, with_text=T > 0)
    inp = make_inputs(geom, 200 + K, rows=3, kv_tokens=L, num_query=K, num_text=T)
    ref_hidden, ref_comp = run_reference(geom, sd, inp, K)
    hid = oracle.qformer_forward(sd, geom, inp["query_embeds"], inp["enc"], inp["input_ids"])
    comp = oracle.proj_norm(sd, hid, K)
    assert np.abs(hid.numpy() - ref_hidden).max() <= 2e-5
    assert np.abs(comp.numpy() - ref_comp).max() <= 2e-6


def test_reference_masks_are_zero_on_the_tdc_path():
    """All-ones attention masks (cambrian_arch.py:1648-1650) => additive masks are exactly 0,
    which is why neither the oracle nor the kernels materialise them."""
    geom = QFormerGeometry(hidden=64, heads=1, intermediate=64, layers=1, cross_freq=1, d_enc=32, d_out=32,
                           vocab=16, max_pos=8)
    model = ref_shim.build_reference_bert(geom, 4)
    ones = torch.ones(2, 6)
    ext = model.get_extended_attention_mask(ones, (2, 6), torch.device("cpu"), False)
    inv = model.invert_attention_mask(ones)
    assert float(ext.abs().max()) == 0.0 and float(inv.abs().max()) == 0.0
