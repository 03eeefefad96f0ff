"""bench.py host-side contract: the algorithmic-FLOP formula equals BASELINE.md §3, the workloads are the
BASELINE configs, and the reference arm prints the required JSON keys (CPU only, tiny sample)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


@pytest.mark.parametrize("L,d_enc,K,T,d_out,gflop", [
    (156, 3072, 16, 0, 3072, 11.91), (206, 3072, 16, 0, 3072, 14.75), (156, 3584, 16, 0, 3584, 13.39),
    (206, 3584, 16, 0, 3584, 16.71), (194, 1152, 16, 0, 3072, 7.16), (626, 1152, 16, 0, 3072, 16.46),
    (1000, 1152, 16, 0, 3072, 24.51), (206, 3584, 64, 0, 3584, 26.13), (206, 3584, 16, 32, 3584, 22.22),
])
def test_flops_per_row_matches_baseline_table(L, d_enc, K, T, d_out, gflop):
    total, kv = bench.flops_per_row(L, d_enc, K, T, d_out)
    assert abs(total / 1e9 - gflop) < 0.02 * gflop      # BASELINE.md quotes 2-4 significant digits
    assert kv == 6 * 2 * 2 * L * d_enc * 768


def test_variant_workloads_and_projector_flops():
    """SURVEY 8.0 / 8d readings exist as --workload variants; the GELU-MLP projector flavour replaces
    vision_proj's 2*K*768*d by 2*K*(768*d + d*d) in the per-row FLOPs (MLP projector formula of BASELINE.md)."""
    lit, seg, ev = (bench.WORKLOADS[k] for k in ("literal_d1152_mlp", "segment_kv_d1152", "eval64x600"))
    assert lit["d_enc"] == 1152 and lit["kv_tokens"] == 194 and lit["projector"] == "gelu_mlp"
    assert seg["kv_tokens"] == 4 * 144 + 50 and seg["d_enc"] == 1152
    assert ev["segments"] * 8 == 64 * 600
    a, kv_a = bench.flops_per_row(194, 1152, 16, 0, 3584, "vision_proj")
    b, kv_b = bench.flops_per_row(194, 1152, 16, 0, 3584, "gelu_mlp")
    assert kv_a == kv_b and b - a == 2 * 16 * 3584 * 3584


def test_one_fps_chunk8_workload_is_the_same_hour():
    # SURVEY.md 8d config 3, reference-faithful variant: 3600 frames at 1 fps in chunks of 8 -> 3150 rows, and the
    # metric's unit stays video-seconds (8 per chunk)
    w = bench.WORKLOADS["hour_1fps_chunk8"]
    assert w["segments"] * w["frames_per_segment"] == 3600
    assert w["segments"] * (w["frames_per_segment"] - 1) == 3150
    assert w["segments"] * bench.vsec(w) == 3600 and bench.vsec(bench.WORKLOADS["hour_qwen7b"]) == 1


def test_default_workload_is_the_north_star_config():
    w = bench.WORKLOADS["hour_qwen7b"]
    assert w["segments"] == 3600 and w["d_enc"] == 3584 and w["d_out"] == 3584 and w["num_query"] == 16
    assert w["segments"] * (w["frames_per_segment"] - 1) == 10800
    assert bench.WORKLOADS["cfg2_llama3b"]["segments"] == 256


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-rows", "2", "--workload", "cfg2_llama3b"], capture_output=True, text=True,
                         timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
