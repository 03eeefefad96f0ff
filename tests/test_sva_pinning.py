"""Pin the SVA oracle (oracle/sva_oracle.py) against the reference's own modules, loaded unmodified
(tdc/vision_sampler.py imports only torch/numpy) and against the real window rearrangement.  Build container only."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_shim, sva_oracle
from oracle.synth import make_sva_state_dict

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


def _load_vision_sampler():
    name = "_tdc_reference_vision_sampler"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref_shim.REFERENCE_ROOT, "tdc", "vision_sampler.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("hidden,sides,layers,sizes", [
    (128, (2, 2), 2, [(384, 384), (384, 384)]),
    (128, (2, 1), 1, [(640, 360), (384, 384), (300, 500)]),     # letterboxed / pillarboxed frames -> real masks
    (256, (2, 2), 3, [(1280, 720)]),
])
def test_token_sampler_and_projector_equal_reference(hidden, sides, layers, sizes):
    vs = _load_vision_sampler()
    tower_dims = (96, 64)
    Q = 4                                             # query grid side (12 in the shipped config)
    sd = make_sva_state_dict(hidden, tower_dims, sides, layers, seed=hidden + layers, stress=3.0)
    bs = len(sizes)
    rs = np.random.RandomState(1)
    tower = [torch.from_numpy(rs.standard_normal((bs, (Q * s) ** 2, c)).astype(np.float32))
             for s, c in zip(sides, tower_dims)]

    # --- reference modules
    sampler = vs.VisionTokenSampler(hidden, hidden, [hidden] * 2, list(sides), hidden, layers).eval()
    own = {k[len("vision_sampler_0."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("vision_sampler_0.")}
    sampler.load_state_dict(own, strict=True)
    proj = []
    for t, c in enumerate(tower_dims):
        m = torch.nn.Sequential(torch.nn.Linear(c, hidden), torch.nn.GELU(), torch.nn.Linear(hidden, hidden),
                                torch.nn.LayerNorm(hidden)).eval()
        m.load_state_dict({k[len(f"mm_projector_aux_{t}."):]: torch.from_numpy(v) for k, v in sd.items()
                           if k.startswith(f"mm_projector_aux_{t}.")}, strict=True)
        proj.append(m)
    from oracle import harness
    arch = harness._load_cambrian_arch()

    class Bare(arch.CambrianMetaForCausalLM):
        def get_model(self):
            return None

    with torch.no_grad():
        feats = [proj[t](tower[t]) for t in range(2)]
        lat, masks = Bare().rearrange_vision_tower_features_inference(feats, Q, sizes)
        nq = Q * Q
        ctx = feats[0].mean(1).view(bs, 1, 1, -1).expand(-1, nq, 1, -1).flatten(0, 1)
        qry = torch.from_numpy(sd["vision_query"])[0].view(1, 1, 1, -1).expand(bs, nq, -1, -1).flatten(0, 1)
        masks_r = [m.view(m.shape[0], 1, 1, -1).expand(-1, -1, 1, -1) for m in masks]   # as VisionCrossAttentionLayer does
        ref = sampler(qry, ctx, *lat, *masks).view(bs, nq, hidden)

    # pieces
    for t in range(2):
        assert torch.allclose(sva_oracle.mm_projector_aux(sd, f"mm_projector_aux_{t}", tower[t]), feats[t], atol=1e-5)
        assert torch.equal(sva_oracle.rearrange_windows(feats[t], Q), lat[t])
        grid = Q * sides[t]
        m = torch.cat([sva_oracle.window_masks(sizes[b], grid, Q) for b in range(bs)], 0)
        assert torch.equal(m, masks[t])
    got = sva_oracle.sva_frames(sd, tower, sizes, Q, layers)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 5e-5


def test_two_query_groups_equal_reference_modules():
    """query_num_list with two groups (cambrian_arch.py:1017-1148): group 1's coarser grid goes through its own
    VisionTokenSampler over larger windows, is resized with F.interpolate(bilinear, align_corners=False) to the final
    grid and concatenated on the feature axis.  Reference modules + the literal interpolate lines vs the oracle."""
    import torch.nn.functional as F
    from oracle.synth import add_sva_group
    vs = _load_vision_sampler()
    hidden, dims, layers, sizes = 128, (96, 64), 2, [(640, 360), (384, 384)]
    final, coarse = 4, 2                               # tower grids 8 x 8: windows 2 x 2 (final) and 4 x 4 (coarse)
    sd = make_sva_state_dict(hidden, dims, (2, 2), layers, seed=5, stress=2.0)
    add_sva_group(sd, 1, hidden, (4, 4), layers, seed=6, stress=2.0)
    bs = len(sizes)
    rs = np.random.RandomState(2)
    tower = [torch.from_numpy(rs.standard_normal((bs, 64, c)).astype(np.float32)) for c in dims]
    from oracle import harness
    arch = harness._load_cambrian_arch()

    class Bare(arch.CambrianMetaForCausalLM):
        def get_model(self):
            return None

    with torch.no_grad():
        feats = [sva_oracle.mm_projector_aux(sd, f"mm_projector_aux_{t}", tower[t]) for t in range(2)]
        outs = []
        for g, (q, sides) in enumerate(((final, (2, 2)), (coarse, (4, 4)))):
            sampler = vs.VisionTokenSampler(hidden, hidden, [hidden] * 2, list(sides), hidden, layers).eval()
            sampler.load_state_dict({k[len(f"vision_sampler_{g}."):]: torch.from_numpy(v) for k, v in sd.items()
                                     if k.startswith(f"vision_sampler_{g}.")}, strict=True)
            lat, masks = Bare().rearrange_vision_tower_features_inference(feats, q, sizes)
            nq = q * q
            ctx = feats[0].mean(1).view(bs, 1, 1, -1).expand(-1, nq, 1, -1).flatten(0, 1)
            qry = torch.from_numpy(sd["vision_query"])[g].view(1, 1, 1, -1).expand(bs, nq, -1, -1).flatten(0, 1)
            o = sampler(qry, ctx, *lat, *masks).view(bs, nq, hidden)
            if q != final:                                                     # cambrian_arch.py:1107-1131
                o = o.permute(0, 2, 1).contiguous().view(bs, -1, q, q)
                o = F.interpolate(o.float(), size=(final, final), mode="bilinear", align_corners=False)
                o = o.permute(0, 2, 3, 1).contiguous().flatten(1, 2)
            outs.append(o)
        ref = torch.cat(outs, -1)
    got = sva_oracle.sva_frames_groups(sd, tower, sizes, (final, coarse), final, layers)   # 16 heads, as the reference
    assert got.shape == ref.shape == (bs, final * final, 2 * hidden)
    assert float((got - ref).abs().max()) <= 5e-5


@pytest.mark.parametrize("sides", [(2, 2), (2, 1)])
def test_sep_layers_equal_reference_modules(sides):
    """VisionTokenSampler(layer_type="sep") = VisionAggregationLayer (vision_sampler.py:404-517): per-tower attention
    (an MLP where the window is a single token) mixed by softmax(weight_mlp): reference module vs the oracle."""
    from oracle.synth import make_sva_sep_state_dict
    vs = _load_vision_sampler()
    hidden, dims, layers, Q = 128, (96, 64), 2, 4
    sizes = [(640, 360), (384, 384), (300, 500)]
    sd = make_sva_sep_state_dict(hidden, dims, sides, layers, seed=11, stress=2.0)
    bs = len(sizes)
    rs = np.random.RandomState(3)
    tower = [torch.from_numpy(rs.standard_normal((bs, (Q * s) ** 2, c)).astype(np.float32)) for s, c in zip(sides, dims)]
    sampler = vs.VisionTokenSampler(hidden, hidden, [hidden] * 2, list(sides), hidden, layers, "sep").eval()
    sampler.load_state_dict({k[len("vision_sampler_0."):]: torch.from_numpy(v) for k, v in sd.items()
                             if k.startswith("vision_sampler_0.")}, strict=True)
    from oracle import harness
    arch = harness._load_cambrian_arch()

    class Bare(arch.CambrianMetaForCausalLM):
        def get_model(self):
            return None

    with torch.no_grad():
        feats = [sva_oracle.mm_projector_aux(sd, f"mm_projector_aux_{t}", tower[t]) for t in range(2)]
        lat, masks = Bare().rearrange_vision_tower_features_inference(feats, Q, sizes)
        nq = Q * Q
        ctx = feats[0].mean(1).view(bs, 1, 1, -1).expand(-1, nq, 1, -1).flatten(0, 1)
        qry = torch.from_numpy(sd["vision_query"])[0].view(1, 1, 1, -1).expand(bs, nq, -1, -1).flatten(0, 1)
        ref = sampler(qry, ctx, *lat, *masks).view(bs, nq, hidden)
    got = sva_oracle.sva_frames(sd, tower, sizes, Q, layers, layer_type="sep")
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 5e-5
