"""TEST INFRASTRUCTURE — the synthetic weight / input generators live in `tdc_video_b200/synth.py` (bench.py's
product arm must not import `oracle/`); re-exported here for the oracle, the golden scripts and the tests."""
from tdc_video_b200.synth import (  # noqa: F401
    QFormerGeometry, make_state_dict, make_inputs, make_sva_state_dict, make_frontend_state_dict, add_sva_group, make_sva_sep_state_dict,
)
