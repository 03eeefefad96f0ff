"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference Q-Former for pinning the oracle.

Loads /root/reference/tdc/Qformer.py as-is (no edits, no copies) under the transformers
version installed here (5.x; the reference pins 4.46).  Three names the file imports from
`transformers.modeling_utils` moved to `transformers.pytorch_utils`; `init_weights()` became
`post_init()`; `get_head_mask` was removed.  The shim restores those names *around* the
import — the reference source itself is executed untouched.

Only available where /root/reference exists (the build container).  Nothing under tests
marked `gpu`, `bench.py` or `smoke()` may import this module (the GPU box has no reference);
they use the restatement in `oracle/qformer_oracle.py`, which `tests/test_oracle_pinning.py`
and `oracle/make_golden.py` pin against this module.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TDC_REFERENCE_ROOT", "/root/reference")
_QFORMER_PATH = os.path.join(REFERENCE_ROOT, "tdc", "Qformer.py")
_MODULE_NAME = "_tdc_reference_qformer"


def reference_available() -> bool:
    return os.path.isfile(_QFORMER_PATH)


def load_reference_qformer() -> types.ModuleType:
    """Return the reference `tdc/Qformer.py` module object (cached in sys.modules)."""
    if _MODULE_NAME in sys.modules:
        return sys.modules[_MODULE_NAME]
    if not reference_available():
        raise FileNotFoundError(f"reference not found at {_QFORMER_PATH}")
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    for name in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        def find_pruneable_heads_and_indices(heads, n_heads, head_size, already_pruned_heads):  # never called on the TDC path
            raise NotImplementedError("head pruning is not part of the TDC path")
        mu.find_pruneable_heads_and_indices = find_pruneable_heads_and_indices

    spec = importlib.util.spec_from_file_location(_MODULE_NAME, _QFORMER_PATH)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_MODULE_NAME] = mod
    spec.loader.exec_module(mod)

    base = mod.BertPreTrainedModel
    if not hasattr(base, "get_head_mask"):
        # head_mask is None at every TDC call site -> per-layer None (what 4.46's get_head_mask returns)
        base.get_head_mask = lambda self, head_mask, num_hidden_layers, *a, **k: [None] * num_hidden_layers
    orig_init_weights = base.init_weights

    def init_weights(self):
        # transformers 5 requires post_init() to have run before init_weights(); the reference
        # calls the 4.x-style self.init_weights() at the end of __init__ (Qformer.py:697, 979).
        if not getattr(self, "_tdc_post_init_done", False):
            self._tdc_post_init_done = True
            return self.post_init()
        return orig_init_weights(self)

    base.init_weights = init_weights
    return mod


def build_reference_bert(geom, num_query: int):
    """Construct the reference `BertModel` for a `QFormerGeometry` (fp32, eval mode)."""
    from transformers.models.bert.configuration_bert import BertConfig

    mod = load_reference_qformer()
    cfg = BertConfig(
        vocab_size=max(geom.vocab, 1),
        hidden_size=geom.hidden,
        num_hidden_layers=geom.layers,
        num_attention_heads=geom.heads,
        intermediate_size=geom.intermediate,
        max_position_embeddings=max(geom.max_pos, 1),
        layer_norm_eps=geom.ln_eps,
        hidden_act="gelu",
    )
    # the attributes the reference adds in cambrian_arch.py:407-412 (init_Qformer)
    cfg.encoder_width = geom.d_enc
    cfg.add_cross_attention = True
    cfg.cross_attention_freq = geom.cross_freq
    cfg.query_length = num_query
    model = mod.BertModel(cfg, add_pooling_layer=False)
    model.eval()
    return model
