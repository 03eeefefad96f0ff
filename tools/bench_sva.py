"""Measure the SVA connector at the shipped geometry: frames/s, algorithmic TFLOP/s against the
measured bf16 peak, and the reference algorithm (oracle port, fp32 torch) on the host cores.  One JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sva_oracle  # noqa: E402  (checker + CPU baseline only)
from tdc_video_b200.synth import make_sva_state_dict  # noqa: E402
from tdc_video_b200.sva import SVAConnector  # noqa: E402


def flops_per_frame(hidden, dims, sides, layers, Q):
    nq = Q * Q
    f = 0
    for c, s in zip(dims, sides):
        n = (Q * s) ** 2
        f += 2 * n * (c * hidden + hidden * hidden)                       # mm_projector_aux
        f += layers * 2 * (2 * n * hidden * hidden)                       # k_proj + v_proj per layer
    per_q = 2 * hidden * hidden * (2 + 1 + 1 + 2)                         # proj_in(2H) + q_proj + o_proj + MLP(2)
    f += layers * (nq * per_q + 2 * hidden * hidden)                      # + proj_context (one vector per frame)
    f += layers * nq * 4 * sum(s * s for s in sides) * hidden             # QK^T + PV
    return f


def main():
    hidden, dims, sides, layers, Q = 1024, (1152, 1536), (2, 2), 3, 12
    frames = int(os.environ.get("SVA_FRAMES", "224"))
    dev = torch.device("cuda", 0)
    sd = make_sva_state_dict(hidden, dims, sides, layers, seed=1, stress=1.0)
    mod = SVAConnector(dims, sides, hidden=hidden, query_side=Q, num_layers=layers)
    mod.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    mod = mod.cuda().eval()
    g = torch.Generator(device=dev).manual_seed(0)
    tower = [torch.randn((frames, (Q * s) ** 2, c), generator=g, device=dev).bfloat16() for s, c in zip(sides, dims)]
    sizes = [(1280, 720)] * frames
    out = mod(tower, sizes)
    torch.cuda.synchronize()
    ref = sva_oracle.sva_frames(sd, [t[:2].float().cpu() for t in tower], sizes[:2], Q, layers)
    from oracle.qformer_oracle import parity_metrics
    pm = parity_metrics(out[:2].float().cpu(), ref)
    for _ in range(3):
        mod(tower, sizes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        mod(tower, sizes)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = flops_per_frame(hidden, dims, sides, layers, Q) * frames
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1400.0}
    torch.set_num_threads(os.cpu_count())
    n_cpu = 4
    t0 = time.perf_counter()
    sva_oracle.sva_frames(sd, [t[:n_cpu].float().cpu() for t in tower], sizes[:n_cpu], Q, layers)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({
        "metric": "SVA connector frames/s (SigLIP 24x24x1152 + DINOv2 24x24x1536 -> 144 x 1024 queries, 3 layers)",
        "value": frames / (ms * 1e-3), "ms": ms, "frames": frames, "parity_first_frames": pm,
        "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops_sustained"],
                     "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                     "gflop_per_frame": fl / frames / 1e9},
        "cpu_baseline": {"value": n_cpu / cpu_s, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{n_cpu} frames, fp32 torch"}}))


if __name__ == "__main__":
    main()
