"""TEST INFRASTRUCTURE ONLY — drive the reference's REAL `prepare_inputs_labels_for_multimodal`
(tdc/cambrian_arch.py:864-1844, unmodified) through a fake model, to pin the driver oracle.

Only usable where /root/reference exists.  The fake model supplies exactly the attributes the
function touches: two stub vision towers returning deterministic [N,144,C]
features, a linear `mm_projector`, `image_newline`, `frame_seg`, the reference's own
`BertLMHeadModel` as `Qformer`, `vision_proj` / `query_proj` / `query_tokens`, an `embed_tokens`
table and a stub BERT tokenizer.  Everything is fp32 on the CPU.

`run_reference_driver(...)` returns the visual token sequence the reference splices into the
prompt together with the inputs of the TDC block (projected frames with newline tokens,
segment boundaries) so that `oracle.driver_oracle.compress_video` can be compared on the very
same data.
"""
from __future__ import annotations

import importlib
import sys
import types
from types import SimpleNamespace

import torch
from torch import nn

from . import ref_shim

IMAGE_TOKEN_INDEX = -200


def _load_cambrian_arch():
    ref_shim.load_reference_qformer()  # installs the transformers-5 compatibility names
    if "IPython" not in sys.modules:  # cambrian_arch.py:45 `from IPython import embed` (debug leftover)
        stub = types.ModuleType("IPython")
        stub.embed = lambda *a, **k: None
        sys.modules["IPython"] = stub
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    try:
        return importlib.import_module("tdc.cambrian_arch")
    finally:
        pass


class _StubTower(nn.Module):
    """Deterministic per-frame features: frame f -> seeded [144, C] pattern (no pixels needed)."""

    def __init__(self, width, table):
        super().__init__()
        self.width = width
        self.table = table  # [n_total_frames, 144, C]

    def forward(self, images):
        # `images` carries the frame ids in its first element of each frame (see run_reference_driver)
        idx = images.reshape(images.shape[0], -1)[:, 0].long()
        return self.table[idx]


class _Tokens:
    def __init__(self, ids):
        self.input_ids = ids

    def to(self, device):
        return self


class _StubBeats:
    """Stands in for BEATs.extract_features (upstream model, out of scope): returns the pre-computed
    `[1, 50 * seconds, 768]` features of the window whose id is stored in the first wav sample."""

    def __init__(self, windows):
        self.windows = windows
        self.calls = 0

    def extract_features(self, wav, padding_mask=None, feature_only=True):
        w = int(round(float(wav[0, 0])))
        self.calls += 1
        return self.windows[w], None


def run_reference_driver(weights, geom, n_frames, *, d_llm, context_token_num=16, query_type="Avg_pool",
                         text_input=True, add_static=True, tokenizer_model_max_length=100000, prompt_ids=None,
                         dino_table=None, siglip_table=None, seed=0, audio_windows=None, video_indices=None,
                         audio_seconds=None):
    """weights: Q-Former state dict (keys relative to `Qformer.bert.`) + `vision_proj.*`, `query_proj.*`,
    `frame_seg`, `query_tokens`, `mm_projector.weight/bias`, `image_newline`, `embed_tokens`.
    Returns dict(visual_tokens, frames, segment_frame_indices, split_sizes)."""
    arch = _load_cambrian_arch()
    qmod = ref_shim.load_reference_qformer()
    from transformers.models.bert.configuration_bert import BertConfig

    t = lambda x: torch.as_tensor(x, dtype=torch.float32)
    c1, c2 = siglip_table.shape[-1], dino_table.shape[-1]

    # --- the reference's own BertLMHeadModel as model.Qformer (cambrian_arch.py:403-424)
    cfg = BertConfig(vocab_size=max(geom.vocab, 1), hidden_size=geom.hidden, num_hidden_layers=geom.layers,
                     num_attention_heads=geom.heads, intermediate_size=geom.intermediate,
                     max_position_embeddings=max(geom.max_pos, 1), layer_norm_eps=geom.ln_eps)
    cfg.encoder_width = d_llm
    cfg.add_cross_attention = True
    cfg.cross_attention_freq = geom.cross_freq
    cfg.query_length = context_token_num
    qformer = qmod.BertLMHeadModel(cfg).eval()
    sd = {"bert." + k: t(v) for k, v in weights.items() if k.startswith(("embeddings.", "encoder."))}
    missing, unexpected = qformer.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("cls.") or "position_ids" in k or (not text_input and (
        ".intermediate.dense" in k or ".output.dense" in k or ".output.LayerNorm" in k or "embeddings.word" in k
        or "embeddings.position" in k)) for k in missing), missing

    def linear(prefix, i, o):
        m = nn.Linear(i, o)
        m.weight.data, m.bias.data = t(weights[prefix + ".weight"]), t(weights[prefix + ".bias"])
        return m

    inner = SimpleNamespace()
    inner.config = SimpleNamespace(
        model_type="qwen2", image_token_len=144, query_num_list=[144], mm_projector_type="mlp2x_gelu",
        tokenizer_model_max_length=tokenizer_model_max_length, context_token_num=context_token_num,
        audio_input=audio_windows is not None, add_static=add_static, text_input=text_input, query_type=query_type,
        max_num_segments=24, lowres_token=8, tokenizer_padding_side="right", hidden_size=d_llm)
    towers = [_StubTower(c1, t(siglip_table)), _StubTower(c2, t(dino_table))]
    inner.get_vision_tower_aux_list = lambda: towers
    if "mm_projector.0.weight" in weights:
        # the shipped projector: Linear -> GELU -> Linear (multimodal_projector/builder.py:40-47, cambrian_arch.py:65-69)
        inner.mm_projector = nn.Sequential(linear("mm_projector.0", c1 + c2, d_llm), nn.GELU(),
                                           linear("mm_projector.2", d_llm, d_llm))
    else:
        inner.mm_projector = linear("mm_projector", c1 + c2, d_llm)
    inner.Qformer = qformer
    inner.query_tokens = nn.Parameter(t(weights["query_tokens"]))
    inner.vision_proj = linear("vision_proj", geom.hidden, d_llm)
    inner.query_proj = linear("query_proj", d_llm, geom.hidden)
    inner.image_newline = nn.Parameter(t(weights["image_newline"]))
    inner.frame_seg = nn.Parameter(t(weights["frame_seg"]))
    emb = nn.Embedding.from_pretrained(t(weights["embed_tokens"]))
    inner.embed_tokens = emb
    ids = torch.as_tensor(prompt_ids if prompt_ids is not None else [[5, 6, 7]], dtype=torch.long).reshape(1, -1)
    inner.bert_tokenizer = lambda prompt, **kw: _Tokens(ids)
    audios = None
    if audio_windows is not None:
        inner.dtype = torch.float32
        inner.audio_proj = linear("audio_proj", 768, d_llm)
        inner.audio_encoder = SimpleNamespace(beats_path="stub", beats=_StubBeats([t(a) for a in audio_windows]))
        secs = int(audio_seconds)
        wav = torch.zeros(1, secs * 16000)
        for w in range(len(audio_windows)):          # window id in the first sample of every 10-s window
            wav[0, w * 10 * 16000] = float(w)
        audios = [{"audio_wav": wav, "audio_wav_mask": torch.zeros(1, secs * 16000, dtype=torch.bool)}]

    class Fake(arch.CambrianMetaForCausalLM, nn.Module):
        def __init__(self):
            nn.Module.__init__(self)
            self.model = inner
            self.config = SimpleNamespace(tokenizer_model_max_length=tokenizer_model_max_length, highres=False,
                                          frame_pos=False)

        def get_model(self):
            return inner

        @property
        def device(self):
            return torch.device("cpu")

    fake = Fake().eval()
    # "images": the stub towers read the frame id from the first element of every frame
    frame_ids = torch.arange(n_frames, dtype=torch.float32)
    img = frame_ids.view(n_frames, 1, 1, 1).expand(n_frames, 3, 2, 2).contiguous()
    images = [[img], [img.clone()]]          # [siglip list, dino list], one video
    input_ids = torch.tensor([[1, IMAGE_TOKEN_INDEX, 2, 3]], dtype=torch.long)
    with torch.no_grad():
        out = fake.prepare_inputs_labels_for_multimodal(
            input_ids, None, None, None, None, images, image_aux_attention_masks_list=None,
            image_sizes=[(384, 384)], video_indices=[None] if video_indices is None else [video_indices],
            prompts=["what happens?"], audios=audios)
        new_input_embeds = out[4]
        visual = new_input_embeds[0, 1:-2]   # between embed(1) and embed(2), embed(3)

        # --- the inputs the TDC block saw, recomputed with the same modules (cambrian_arch.py:946-1299)
        dino = towers[1](img)
        _, split_sizes, _, _, seg_idx_all = fake.adapt_segment(dino, [n_frames], [img, img], max_num_segments=24)
        sel = torch.arange(n_frames)
        if n_frames > 224:
            interval = n_frames / 224.0
            sel = torch.tensor([int(interval * i) for i in range(224)])
        feats = torch.cat([towers[0](img[sel]), towers[1](img[sel])], -1)
        proj = inner.mm_projector(feats).view(len(sel), 12, 12, -1)
        nl = inner.image_newline.view(1, 1, 1, -1).expand(len(sel), 12, 1, -1)
        frames = torch.cat([proj, nl], dim=2).flatten(1, 2)       # [n, 156, d]
    return dict(visual_tokens=visual, frames=frames, segment_frame_indices=seg_idx_all[0], split_sizes=split_sizes,
                prompt_ids=ids)
