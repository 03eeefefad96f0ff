#!/usr/bin/env python
"""bench.py — TDC compression throughput (video-seconds/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hour_qwen7b|cfg2_llama3b|...] [--num-text T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

One "step" = one pass of the hot path over one synthetic video per GPU, in the reference's order
(tdc/cambrian_arch.py): mm_projector on every frame's tower features (:1149), image_newline (:1269-1281),
audio_proj (:1611-1614), query build from the key frame (:1629-1640), the Q-Former on every dynamic frame
(:1653-1662), vision_proj + L2-normalise (:1664-1667) — for all chunks at once through ONE library entry
(`tdc_compress_frames`, --entry frames, the default) — then (N > 1) the all-gather of the compressed tokens
so that every rank holds the ordered sequence (fused into the final kernel through NVSwitch multicast stores;
`--no-multicast` = NCCL all-gather).

 * `value`  : whole-job video-seconds/s, the towers' outputs resident in HBM, CUDA-event timed, max over ranks
 * `e2e`    : the same through `QFormerEngine.compress_frames_host` with the towers' outputs in pinned HOST
              memory (H2D of every frame inside the timed region, compressed tokens copied back to the host)
 * `roofline`: dominant kernel = the cross-attention K/V projection GEMM (tcgen05), EXECUTED FLOPs / its
              CUDA-event time (events on the launching stream, recorded by the library)
 * `path`   : model FLOPs (reference formulation) and executed FLOPs (after weight folding) per step
 * `unfolded`: the same step with every projection executed as the reference orders them (fold = 0)
 * `qformer_only`: round 1's line — Q-Former + vision_proj on already projected tokens resident in HBM
              (`--entry tokens` runs only that; the variant workloads use it)
 * `cpu_baseline`: the reference algorithm (oracle port, fp32 torch) on the host cores, bounded sample

Synthetic data, random-init weights (tdc_video_b200/synth.py statistics); weak scaling: every GPU
compresses its own `segments` video-seconds; `strong` = ONE video sharded over the N GPUs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json config 3 / north_star target: 1-hour video, TDC-Qwen2-7B shapes, reference order
    # (frame tokens already in LLM width): 3600 one-second segments x 4 frames, frame 0 static ->
    # 3 rows per segment; row KV = 144 visual + 12 newline + 50 audio tokens.
    "hour_qwen7b": dict(segments=3600, frames_per_segment=4, kv_tokens=206, audio_tokens=50, d_enc=3584, d_out=3584,
                        num_query=16, num_text=0, label="1-hour video, Qwen2-7B widths (d=3584), L=206, K=16"),
    # SURVEY.md 8d config 3, reference-faithful variant: the same hour sampled at 1 fps (3600 frames), chunks of 8 frames
    # (cambrian_arch.py:1603-1628: <= 8 frames per Q-Former call, the first one the key frame) -> 450 chunks x 7 rows
    "hour_1fps_chunk8": dict(segments=450, frames_per_segment=8, video_seconds_per_segment=8, kv_tokens=206,
                             audio_tokens=50, d_enc=3584, d_out=3584, num_query=16, num_text=0,
                             label="1-hour video at 1 fps in 8-frame chunks (450 chunks x 7 rows), Qwen2-7B widths, "
                                   "L=206, K=16"),
    # BASELINE.json config 2: 256 segments, Llama-3.2-3B widths
    "cfg2_llama3b": dict(segments=256, frames_per_segment=4, kv_tokens=206, audio_tokens=50, d_enc=3072, d_out=3072,
                         num_query=16, num_text=0, label="256-segment video, Llama-3.2-3B widths (d=3072), L=206, K=16"),
    # ---- variants of SURVEY.md 8.0 / 8d (not the driver's line; run with --workload) -------------------------------
    # BASELINE-literal order: Q-Former on the towers' own width (SigLIP d=1152, 144 visual + 50 audio tokens), then
    # the GELU-MLP projector 768 -> d_llm instead of vision_proj + normalise
    "literal_d1152_mlp": dict(segments=3600, frames_per_segment=4, kv_tokens=194, audio_tokens=50, d_enc=1152,
                              d_out=3584, num_query=16, num_text=0, projector="gelu_mlp",
                              label="1-hour video, BASELINE-literal order: Q-Former at d_enc=1152, L=194, then "
                                    "GELU-MLP projector 768->3584"),
    # north-star reading "segment KV": every row attends to all F*144 + 50 tokens of its segment
    "segment_kv_d1152": dict(segments=3600, frames_per_segment=4, kv_tokens=626, audio_tokens=50, d_enc=1152,
                             d_out=3584, num_query=16, num_text=0,
                             label="1-hour video, segment-level KV (L = 4*144 + 50 = 626), d_enc=1152"),
    # BASELINE.json config 5: 64 concurrent 10-minute videos = 38 400 segments over 8 GPUs -> 4800 per GPU
    # (K = 16 default; --num-query 64 for the sweep)
    "eval64x600": dict(segments=4800, frames_per_segment=4, kv_tokens=206, audio_tokens=50, d_enc=3584, d_out=3584,
                       num_query=16, num_text=0,
                       label="64 x 10-minute videos over 8 GPUs (4800 segments per GPU), Qwen2-7B widths, L=206"),
}
H, I, LAYERS, HEADS, N_CROSS = 768, 3072, 12, 12, 6


def flops_per_row(L, d_enc, K, T, d_out, projector="vision_proj"):
    """Algorithmic FLOPs of one row, reference formulation."""
    n = K + T
    kv = N_CROSS * 2 * (2 * L * d_enc * H)
    proj = 2 * K * H * d_out if projector == "vision_proj" else 2 * K * (H * d_out + d_out * d_out)
    rest = (LAYERS * (8 * n * H * H + 4 * n * n * H) + N_CROSS * (4 * K * H * H + 4 * K * L * H)
            + LAYERS * 4 * K * H * I + LAYERS * 4 * T * H * I + proj)
    return kv + rest, kv


def vsec(w):
    """video-seconds one segment (= chunk) stands for"""
    return w.get("video_seconds_per_segment", 1)


def frames_flops(w, tv=144, d_in=1024, d_audio=768):
    """Per SEGMENT (= one video-second in the canonical unit) of the frames entry: (model FLOPs in the reference formulation, executed FLOPs with the
    folded weights, executed FLOPs of the K/V projection GEMMs alone)."""
    F_, L, d, K, T = w["frames_per_segment"], w["kv_tokens"], w["d_enc"], w["num_query"], w.get("num_text", 0)
    ta = w["audio_tokens"]
    f_row, f_row_kv = flops_per_row(L, d, K, T, w["d_out"])
    proj_frame = 2 * tv * (d_in * d + d * d)                    # mm_projector, every frame (:1149)
    audio_frame = 2 * ta * d_audio * d                          # audio_proj, every frame (:1613)
    query_chunk = 2 * K * d * H                                 # query_proj, once per chunk (:1638)
    model = F_ * (proj_frame + audio_frame) + query_chunk + (F_ - 1) * f_row
    kv_exec_row = 2 * tv * d * 2 * H * N_CROSS + 2 * ta * d_audio * 2 * H * N_CROSS      # folded K/V GEMMs
    executed = (F_ * 2 * tv * d_in * d                          # mm_projector.0 on every frame
                + 2 * tv * d * d + audio_frame + query_chunk    # key frame: mm_projector.2, audio_proj, query_proj
                + (F_ - 1) * (f_row - f_row_kv + kv_exec_row))
    return model, executed, (F_ - 1) * kv_exec_row


def measured_traffic_per_row(name="kv_gemm_traffic.json"):
    """DRAM bytes per row of the KV-projection GEMM from the committed ncu --set full capture
    (profiles/kv_gemm_traffic*.json: dram__bytes_read.sum + dram__bytes_write.sum over the rows of that launch)."""
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["rows_in_launch"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops_sustained=d["bf16_tflops_sustained"], tflops_burst=d["bf16_tflops"], hbm=d["hbm_gbs"],
                    source="MEASURED_PEAKS.json")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples under load = upper half of the clock samples (idle samples bracket the region)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores nearest to GPU `index` (NVML affinity) so that the pinned host
    buffers of the e2e leg are first-touched on the GPU's own NUMA node.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def build_problem(w, seed):
    """Synthetic video of the workload: per-row KV tokens, one query set per segment (Avg_pool-style:
    all rows of a segment share the queries derived from its static frame)."""
    from tdc_video_b200.synth import QFormerGeometry, make_state_dict
    T = w.get("num_text", 0)
    geom = QFormerGeometry(d_enc=w["d_enc"], d_out=w["d_out"], vocab=30522 if T else 0)
    sd = make_state_dict(geom, seed, with_text=T > 0)
    if w.get("projector") == "gelu_mlp":   # mm_projector-style Sequential: `0.*` Linear(768 -> d), `2.*` Linear(d -> d)
        rs = np.random.RandomState(seed + 17)
        d = w["d_out"]
        sd["mm_projector.0.weight"] = (rs.standard_normal((d, geom.hidden)) * 0.02).astype(np.float32)
        sd["mm_projector.0.bias"] = (rs.standard_normal(d) * 0.02).astype(np.float32)
        sd["mm_projector.2.weight"] = (rs.standard_normal((d, d)) * 0.02).astype(np.float32)
        sd["mm_projector.2.bias"] = (rs.standard_normal(d) * 0.02).astype(np.float32)
    if w.get("entry") == "frames":
        from tdc_video_b200.synth import make_frontend_state_dict
        sd.update(make_frontend_state_dict(w["d_enc"], w["d_frame_in"], 768 if w["audio_tokens"] else 0, geom.hidden,
                                           seed + 29, w["num_query"]))
    rows = w["segments"] * (w["frames_per_segment"] - 1)
    return geom, sd, rows


def cpu_baseline_frames(geom, sd, w, sample_rows, seed, passes=1):
    """The reference algorithm of the frames entry (oracle/frames_oracle.py: mm_projector, newline, audio_proj,
    query build, Q-Former, vision_proj + normalise), fp32 torch on all host cores, all sampled rows in ONE call."""
    from oracle import frames_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    F_ = w["frames_per_segment"]
    segs = max(1, sample_rows // (F_ - 1))
    rs = np.random.RandomState(seed)
    frames = torch.from_numpy(rs.standard_normal((segs * F_, 144, w["d_frame_in"])).astype(np.float32))
    audio = None
    if w["audio_tokens"]:
        audio = torch.from_numpy((rs.standard_normal((segs * F_, w["audio_tokens"], 768)) * 0.5).astype(np.float32))
    T = w.get("num_text", 0)
    ids = None if T == 0 else torch.from_numpy(rs.randint(1000, 30000, size=(1, T)))
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    cs, cl = np.arange(segs) * F_, np.full(segs, F_)
    with torch.no_grad():
        frames_oracle.frames_stage(sd_t, geom, frames[:F_], None if audio is None else audio[:F_], cs[:1], cl[:1],
                                   w["num_query"], ids)                                   # warm-up
        t0 = time.perf_counter()
        for _ in range(passes):
            frames_oracle.frames_stage(sd_t, geom, frames, audio, cs, cl, w["num_query"], ids)
        dt = time.perf_counter() - t0
    return passes * segs * vsec(w) / dt, dt, cores


def cpu_baseline(geom, sd, w, sample_rows, seed, passes=1):
    if w.get("entry") == "frames":
        return cpu_baseline_frames(geom, sd, w, sample_rows, seed, passes)
    return cpu_baseline_tokens(geom, sd, w, sample_rows, seed, passes)


def cpu_baseline_tokens(geom, sd, w, sample_rows, seed, passes=1):
    """The reference algorithm (oracle port of tdc/Qformer.py + vision_proj + normalize), fp32 torch on
    all host cores, batched as ONE call (kinder to the CPU than the reference's <= 7-row loop)."""
    from oracle import qformer_oracle as oracle
    from tdc_video_b200.synth import make_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T = w.get("num_text", 0)
    inp = make_inputs(geom, seed, sample_rows, w["kv_tokens"], w["num_query"], T, audio_tokens=w["audio_tokens"])
    ids = None if T == 0 else np.repeat(inp["input_ids"][:1], sample_rows, axis=0)   # one prompt for the whole video
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    if w.get("projector") == "gelu_mlp":
        def run(q, e, i):
            h = oracle.qformer_forward(sd_t, geom, q, e, i)[:, :w["num_query"]]
            return oracle.gelu_mlp(sd_t["mm_projector.0.weight"], sd_t["mm_projector.0.bias"],
                                   sd_t["mm_projector.2.weight"], sd_t["mm_projector.2.bias"], h)
    else:
        def run(q, e, i):
            return oracle.compress(sd_t, geom, q, e, i)
    with torch.no_grad():
        run(inp["query_embeds"][:2], inp["enc"][:2], None if ids is None else ids[:2])  # warm-up
        t0 = time.perf_counter()
        for _ in range(passes):
            run(inp["query_embeds"], inp["enc"], ids)
        dt = time.perf_counter() - t0
    rows_per_s = passes * sample_rows / dt
    return rows_per_s / (w["frames_per_segment"] - 1) * vsec(w), dt, cores


def cpu_baseline_reference_batching(geom, sd, w, sample_rows, seed):
    """The same CPU port called the way the reference's chunk loop calls the Q-Former: <= 7 rows per call
    (tdc/cambrian_arch.py:1603-1692), one pass over the sample."""
    from oracle import qformer_oracle as oracle
    from tdc_video_b200.synth import make_inputs
    T = w.get("num_text", 0)
    inp = make_inputs(geom, seed, sample_rows, w["kv_tokens"], w["num_query"], T, audio_tokens=w["audio_tokens"])
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    ids = None if T == 0 else np.repeat(inp["input_ids"][:1], sample_rows, axis=0)
    with torch.no_grad():
        t0 = time.perf_counter()
        for r0 in range(0, sample_rows, 7):
            sl = slice(r0, min(r0 + 7, sample_rows))
            oracle.compress(sd_t, geom, inp["query_embeds"][sl], inp["enc"][sl], None if ids is None else ids[sl])
        dt = time.perf_counter() - t0
    return sample_rows / dt / (w["frames_per_segment"] - 1) * vsec(w)


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    geom, sd, rows = build_problem(w, 1234)
    sample = args.cpu_sample_rows
    passes = args.cpu_passes or 2
    vals, dts = [], []
    for i in range(args.warmup + args.steps):
        v, dt, cores = cpu_baseline(geom, sd, w, sample, 4321 + i, passes)
        if i >= args.warmup:
            vals.append(v); dts.append(dt)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "video_seconds_per_sec", "value": value, "unit": "video-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(dts),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["label"], "rows_per_step_sample": sample * passes,
                   "entry": "frames (oracle/frames_oracle.py: mm_projector, newline, audio_proj, query build, Q-Former, "
                            "vision_proj)" if w.get("entry") == "frames" else "tokens (oracle/qformer_oracle.py)"},
        "cpu_baseline": {"value": value, "unit": "video-s/s", "cores": cores, "kind": "port",
                         "sample": f"{passes} x {sample} rows (= {passes * sample / (w['frames_per_segment'] - 1) * vsec(w):.1f} "
                                   f"video-s) of the workload per step, oracle port of the reference (fp32 torch, "
                                   f"{sample}-row batches)"},
        "e2e": {"value": value, "unit": "video-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hour_qwen7b", choices=sorted(WORKLOADS))
    ap.add_argument("--segments", type=int, default=0, help="override segments per GPU")
    ap.add_argument("--cpu-sample-rows", type=int, default=96, help="rows per CPU batch (cpu_baseline / reference arm)")
    ap.add_argument("--cpu-passes", type=int, default=0,
                    help="passes over the CPU sample (default: 16 for cpu_baseline = about 10-20 s, 2 per reference-arm step)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--strong-steps", type=int, default=5, help="N > 1: timed steps of the strong-scaling sub-record")
    ap.add_argument("--parity-rows", type=int, default=4,
                    help="rows of the timed workload re-checked against the CPU oracle after the run (0 = skip)")
    ap.add_argument("--e2e-rows-per-batch", type=int, default=600,
                    help="row batch of the host-streaming leg (smaller = shorter pipeline fill/drain; the leg is PCIe-bound)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cta-group", type=int, default=0)
    ap.add_argument("--num-query", type=int, default=0, help="override K (BASELINE config 5 sweeps K = 16 and 64)")
    ap.add_argument("--num-text", type=int, default=0, help="prompt tokens T shared by all rows (text_input mode; default 0 = north-star)")
    ap.add_argument("--gather-batches", type=int, default=2, help="row batches per step at N > 1 (comm/compute overlap)")
    ap.add_argument("--no-multicast", action="store_true", help="N > 1: use the NCCL all-gather instead of multicast stores")
    ap.add_argument("--entry", default="auto", choices=["auto", "frames", "tokens"],
                    help="frames: from the towers' outputs through tdc_compress_frames (default where the workload has "
                         "d_enc == d_out); tokens: Q-Former + vision_proj on already projected tokens (round 1's line)")
    ap.add_argument("--no-fold", action="store_true", help="frames entry: run every projection as the reference orders them")
    ap.add_argument("--unfolded-steps", type=int, default=3, help="frames entry, N = 1: timed steps of the fold = 0 sub-record")
    ap.add_argument("--no-qformer-only", action="store_true", help="skip the round-1 style sub-record (N = 1 only)")
    ap.add_argument("--e2e-chunks-per-batch", type=int, default=600, help="chunks per H2D range of compress_frames_host")
    ap.add_argument("--e2e-no-head-taper", action="store_true", help="dev: first H2D range as large as the others")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.segments:
        w["segments"] = args.segments
    w["num_text"] = args.num_text
    if args.num_query:
        w["num_query"] = args.num_query
    w.setdefault("projector", "vision_proj")
    frames_ok = w["projector"] == "vision_proj" and w["d_enc"] == w["d_out"] and w["kv_tokens"] == 144 + 12 + w["audio_tokens"]
    if args.entry == "frames" and not frames_ok:
        raise SystemExit(f"workload {args.workload} has no frames entry (needs d_enc == d_out and L = 156 + audio)")
    w["entry"] = "frames" if (args.entry == "frames" or (args.entry == "auto" and frames_ok)) else "tokens"
    w["d_frame_in"] = 1024
    if args.impl == "reference":
        return run_reference_arm(args, w)

    import torch.distributed as dist
    from tdc_video_b200 import QFormerEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = dict(world=world, rank=rank, local_rank=local_rank, dev=dev, dist=dist)
    if w["entry"] == "frames":
        line = bench_frames(args, w, ctx)
        if world == 1 and not args.no_qformer_only:
            torch.cuda.empty_cache()
            sub_args = argparse.Namespace(**vars(args))
            sub_args.no_e2e, sub_args.no_cpu_baseline, sub_args.parity_rows = True, True, 0
            q = bench_tokens(sub_args, dict(w, entry="tokens"), ctx)
            line["qformer_only"] = {
                "desc": "round 1's line: Q-Former + vision_proj + L2-normalise on already projected tokens "
                        "[rows, 206, 3584] resident in HBM (tdc_compress)",
                "value": q["value"], "unit": q["unit"], "ms_per_step": q["ms_per_step"], "steps": q["steps"],
                "roofline": q["roofline"], "path": q["path"], "gpu_launches": q["gpu_launches"]}
    else:
        line = bench_tokens(args, w, ctx)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_frames(args, w, ctx):
    """The path from the towers' outputs (tdc_compress_frames).  Returns the JSON line (rank 0; None elsewhere)."""
    from tdc_video_b200 import QFormerEngine
    world, rank, local_rank, dev, dist = ctx["world"], ctx["rank"], ctx["local_rank"], ctx["dev"], ctx["dist"]
    geom, sd, rows = build_problem(w, 1234)
    L, K, d, S, T, F_ = w["kv_tokens"], w["num_query"], w["d_enc"], w["segments"], w["num_text"], w["frames_per_segment"]
    Ta, d_in = w["audio_tokens"], w["d_frame_in"]
    fold = not args.no_fold
    eng = QFormerEngine(d_enc=d, d_out=d, vocab=30522 if T else 0, device=dev, gemm_cta_group=args.cta_group,
                        d_frame_in=d_in, d_audio=768 if Ta else 0)
    eng.load_weights(sd)
    n_frames = S * F_

    # ---- synthetic tower outputs, generated on the device (8e9 normals are too slow on the host), copied to pinned
    # HOST memory — the host copy is the source of truth: the resident tensors are uploaded from it and the e2e leg
    # streams it every step
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    pinned = not args.no_e2e

    def host_empty(shape):
        nonlocal pinned
        if pinned:
            try:
                return torch.empty(shape, dtype=torch.bfloat16, pin_memory=True)
            except RuntimeError as e:
                print(f"[bench] rank {rank}: pinned allocation failed ({e}); using pageable host memory", file=sys.stderr)
                pinned = False
        return torch.empty(shape, dtype=torch.bfloat16)

    frames_host = host_empty((n_frames, 144, d_in))
    audio_host = host_empty((n_frames, Ta, 768)) if Ta else None
    for f0 in range(0, n_frames, 2048):
        f1 = min(n_frames, f0 + 2048)
        frames_host[f0:f1].copy_(torch.randn((f1 - f0, 144, d_in), generator=g, device=dev).to(torch.bfloat16))
        if Ta:
            audio_host[f0:f1].copy_((torch.randn((f1 - f0, Ta, 768), generator=g, device=dev) * 0.5).to(torch.bfloat16))
    frames_dev = frames_host.to(dev)
    audio_dev = audio_host.to(dev) if Ta else None
    ids_dev = None
    if T > 0:   # one BERT-tokenised prompt per video, shared by every row (cambrian_arch.py:1532, 1643-1644)
        ids_dev = torch.randint(1000, 30000, (1, T), generator=torch.Generator().manual_seed(7)).to(dev)

    # ---- integer plan: every segment is one chunk of F frames, frame 0 the key frame (cambrian_arch.py:1603-1628)
    chunk_start = np.arange(S, dtype=np.int64) * F_
    chunk_len = np.full(S, F_, dtype=np.int64)

    def plan_range(c0, c1):
        """index tensors of chunks [c0, c1) — frame indices are global, chunk ids relative to the range"""
        st = torch.from_numpy(chunk_start[c0:c1].astype(np.int32)).to(dev)
        rf = (st[:, None] + torch.arange(1, F_, device=dev, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
        rc = torch.arange(c1 - c0, device=dev, dtype=torch.int32).repeat_interleave(F_ - 1)
        return st, rf, rc

    static_out = torch.empty((S, 144 + 12 + Ta, d), dtype=torch.bfloat16, device=dev)   # key frames pass through
    nb = max(1, args.gather_batches) if world > 1 else 1
    cb = [(S * b // nb, S * (b + 1) // nb) for b in range(nb)]
    plans = [plan_range(c0, c1) for c0, c1 in cb]
    full_plan = plan_range(0, S)
    mcast, exchange = None, "none (1 GPU)"
    gathered = None
    if world > 1:
        exchange = "NCCL all-gather of compressed tokens"
        if not args.no_multicast:
            try:
                from tdc_video_b200.dist import MulticastGather
                mcast = MulticastGather(rows, (K, d), torch.bfloat16, dev)
                exchange = "NVSwitch multicast stores (multimem.st) from the final kernel + device barrier"
            except Exception as e:  # transport fallback only; compute path is identical
                if rank == 0:
                    print(f"[bench] symmetric-memory multicast unavailable ({type(e).__name__}: {e}); using NCCL",
                          file=sys.stderr)
        if mcast is None:
            gathered = torch.empty((world, rows, K, d), dtype=torch.bfloat16, device=dev)

    def compute_range(plan, c0, c1, use_fold=True, mc_ptr=None, want_static=True):
        st, rf, rc = plan
        so, out = eng.compress_frames(frames_dev, st, rf, rc, audio=audio_dev, input_ids=ids_dev, num_query=K,
                                      fold=use_fold, want_static=want_static, out_dtype=torch.bfloat16,
                                      multicast_ptr=mc_ptr, static_into=static_out[c0:c1] if want_static else None)
        return out

    def make_step(chunks, plans_, mc, gath, chunk0=0, use_fold=True):
        """step over the chunk ranges `chunks` (relative to chunk0 for the output slots)"""
        def step():
            if world == 1:
                return compute_range(plans_[0], chunks[0][0], chunks[0][1], use_fold)
            if mc is not None:
                # all-gather fused into the producing kernel: the L2-normalise kernel stores every row through the
                # multicast mapping, so all ranks receive it while the kernel runs
                for (c0, c1), pl in zip(chunks, plans_):
                    compute_range(pl, c0, c1, use_fold, mc_ptr=mc.slot_ptr((c0 - chunk0) * (F_ - 1)))
                mc.barrier()
                return mc.gathered
            works = []
            for (c0, c1), pl in zip(chunks, plans_):
                out = compute_range(pl, c0, c1, use_fold)
                r0, r1 = (c0 - chunk0) * (F_ - 1), (c1 - chunk0) * (F_ - 1)
                works.append(dist.all_gather([gath[w_, r0:r1] for w_ in range(world)], out, async_op=True))
            for wk in works:
                wk.wait()
            return gath.view(-1, K, d)
        return step

    step = make_step(cb, plans, mcast, gathered, 0, fold)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    for _ in range(args.warmup):
        out = step()
    barrier()
    assert torch.isfinite(out[:8].float()).all()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    # per-kernel-class device times come from a SEPARATE profiled pass (two CUDA events around every launch would
    # otherwise sit inside the timed region: ~1800 event records per step)
    psteps = max(1, min(args.steps, 5))
    eng.set_profiling(True)
    eng.reset_profile()
    for _ in range(psteps):
        step()
    barrier()
    prof = eng.profile()
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    mine = torch.tensor([ms / args.steps, sum(v["ms"] for v in prof.values()) / psteps], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine)
    per_rank = {"ms_per_step": [round(float(x[0]), 2) for x in per_rank],
                "kernel_ms_per_step": [round(float(x[1]), 2) for x in per_rank]}
    ms_step = float(t.item()) / args.steps
    value = world * S * vsec(w) / (ms_step * 1e-3)

    # ---- N > 1: the exchange verifies itself, then ONE video sharded over the N GPUs (strong scaling)
    exchange_check, strong = None, None
    if world > 1:
        final = step()
        barrier()
        local_all = compute_range(full_plan, 0, S, fold, want_static=False)
        ref_gather = torch.empty((world * rows, K, d), dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(ref_gather, local_all.contiguous())
        flag = torch.tensor([int(torch.equal(final.reshape(world * rows, K, d), ref_gather))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange_check = bool(flag.item())
        del ref_gather, local_all, final

        S_s = S // world                        # this rank's contiguous range of video-seconds of THE video
        lo = rank * S_s
        # ranges of <= 600 chunks (= one 1800-row internal batch of the library), at least `gather_batches` of them
        nb_s = max(nb, -(-S_s // 600))
        sb = [(lo + S_s * b // nb_s, lo + S_s * (b + 1) // nb_s) for b in range(nb_s)]
        mc_s, gath_s = None, None
        if mcast is not None:
            from tdc_video_b200.dist import MulticastGather
            mc_s = MulticastGather(S_s * (F_ - 1), (K, d), torch.bfloat16, dev)
        else:
            gath_s = torch.empty((world, S_s * (F_ - 1), K, d), dtype=torch.bfloat16, device=dev)
        strong_plans = [plan_range(c0, c1) for c0, c1 in sb]
        strong_step = make_step(sb, strong_plans, mc_s, gath_s, lo, fold)
        strong_ms = timed(strong_step, args.strong_steps)
        n1_ms = timed(lambda: compute_range(full_plan, 0, S, fold), max(2, args.strong_steps // 2))
        mine_s = compute_range(plan_range(lo, lo + S_s), lo, lo + S_s, fold, want_static=False)
        got = strong_step().reshape(world, S_s * (F_ - 1), K, d)[rank]
        flag = torch.tensor([int(torch.equal(got, mine_s))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        # the same step when the key frames' pass-through tokens (produced sharded by the upstream entry: each rank
        # projects its own key frames) are ALSO assembled on every rank: one NCCL all-gather per chunk range, issued
        # right after the range's kernels so that it overlaps the next range's compute
        Ls = 144 + 12 + Ta
        key_bytes = S * Ls * d * 2
        key_all = [torch.empty((world, c1 - c0, Ls, d), dtype=torch.bfloat16, device=dev) for c0, c1 in sb]

        def strong_step_with_keys():
            works = []
            for i, ((c0, c1), pl) in enumerate(zip(sb, strong_plans)):
                compute_range(pl, c0, c1, fold, mc_ptr=None if mc_s is None else mc_s.slot_ptr((c0 - lo) * (F_ - 1)))
                works.append(dist.all_gather_into_tensor(key_all[i].view(world * (c1 - c0), Ls, d), static_out[c0:c1],
                                                         async_op=True))
            if mc_s is not None:
                mc_s.barrier()
            for wk in works:
                wk.wait()

        strong_keys_ms = timed(strong_step_with_keys, args.strong_steps) if mc_s is not None else None
        del key_all
        # ... and with the key frames' tokens stored through a second multicast mapping by the kernel that assembles
        # them (tdc_frames_args.static_multicast): no collective at all, one device barrier closes both exchanges
        strong_keys_mc_ms, keys_match = None, None
        if mc_s is not None:
            mc_k = MulticastGather(S_s, (Ls, d), torch.bfloat16, dev)

            def strong_step_keys_multicast():
                for (c0, c1), (st, rf, rc) in zip(sb, strong_plans):
                    eng.compress_frames(frames_dev, st, rf, rc, audio=audio_dev, input_ids=ids_dev, num_query=K,
                                        fold=fold, out_dtype=torch.bfloat16,
                                        multicast_ptr=mc_s.slot_ptr((c0 - lo) * (F_ - 1)),
                                        static_multicast_ptr=mc_k.slot_ptr(c0 - lo))
                mc_s.barrier()
                mc_k.barrier()

            key_ready = [torch.cuda.Event() for _ in sb]

            def strong_step_keys_overlapped(mode="dma"):
                # the key frames' tokens are complete early in the call (static_ready_event): the copy engines (or a
                # few CTAs) on side streams ship them to every rank while the range's rows are being compressed
                for i, ((c0, c1), (st, rf, rc)) in enumerate(zip(sb, strong_plans)):
                    so, _ = eng.compress_frames(frames_dev, st, rf, rc, audio=audio_dev, input_ids=ids_dev, num_query=K,
                                                fold=fold, out_dtype=torch.bfloat16,
                                                multicast_ptr=mc_s.slot_ptr((c0 - lo) * (F_ - 1)),
                                                static_ready_event=key_ready[i])
                    mc_k.put_async(so, c0 - lo, after=key_ready[i], mode=mode)
                mc_k.barrier()
                mc_s.barrier()

            strong_keys_ov_ms = timed(strong_step_keys_overlapped, args.strong_steps)
            strong_keys_mm_ms = timed(lambda: strong_step_keys_overlapped("multimem"), args.strong_steps)
            compute_range(plan_range(lo, lo + S_s), lo, lo + S_s, fold)
            ref_keys = torch.empty((world * S_s, Ls, d), dtype=torch.bfloat16, device=dev)
            dist.all_gather_into_tensor(ref_keys, static_out[lo:lo + S_s].contiguous())
            mc_k.buf.zero_()
            barrier()
            strong_step_keys_overlapped()
            barrier()
            kflag = torch.tensor([int(torch.equal(mc_k.gathered, ref_keys))], device=dev)
            dist.all_reduce(kflag, op=dist.ReduceOp.MIN)
            keys_ov_match = bool(kflag.item())
            mc_k.buf.zero_()
            barrier()
            strong_keys_mc_ms = timed(strong_step_keys_multicast, args.strong_steps)
            # the multicast buffer must equal an NCCL all-gather of every rank's own key-frame tokens
            strong_step_keys_multicast()
            barrier()
            kflag = torch.tensor([int(torch.equal(mc_k.gathered, ref_keys))], device=dev)
            dist.all_reduce(kflag, op=dist.ReduceOp.MIN)
            keys_match = bool(kflag.item())
            del ref_keys, mc_k
        strong = {"segments_total": S, "segments_per_gpu": S_s, "rows_per_gpu": S_s * (F_ - 1), "ms_per_step": strong_ms,
                  "with_key_frames": None if strong_keys_ms is None else {
                      "ms_per_step": strong_keys_ov_ms, "speedup_vs_n1": n1_ms / strong_keys_ov_ms,
                      "gathered_bytes": key_bytes, "exchange_check": bool(keys_ov_match and keys_match),
                      "multimem_copy_side_stream_ms_per_step": strong_keys_mm_ms,
                      "multicast_from_assemble_kernel_ms_per_step": strong_keys_mc_ms,
                      "nccl_all_gather_ms_per_step": strong_keys_ms,
                      "desc": "every rank also ends with ALL key frames' pass-through tokens [S, 206, d], i.e. the complete "
                              "ordered sequence of the video.  ms_per_step: tdc_compress_frames records "
                              "static_ready_event once the key frames of a range are assembled, and the copy engines "
                              "write them into every peer's symmetric buffer on side streams (tdc_peer_copy; no SM taken "
                              "from the persistent compute kernels) while the range's rows are compressed; "
                              "multimem_copy_side_stream: the same with 16 CTAs storing through a second NVSwitch "
                              "multicast mapping (tdc_multicast_copy) — they have to wait for an SM; "
                              "multicast_from_assemble_kernel: the assembling kernel itself stores through the mapping "
                              "(static_multicast = 1; its stores are fabric-bound and stall the compute stream); "
                              "nccl_all_gather: one NCCL all-gather per chunk range instead.  The synthetic unit has one key "
                              "frame per second (4-frame chunks); the reference's 8-frame chunks halve this payload"},
                  "value": S * vsec(w) / (strong_ms * 1e-3), "unit": "video-s/s", "one_gpu_ms_per_step": n1_ms,
                  "speedup_vs_n1": n1_ms / strong_ms, "steps": args.strong_steps, "exchange": exchange,
                  "own_rows_match": bool(flag.item()),
                  "limiter": "per-GPU step = compute of S/N segments + the device barrier that closes the multicast "
                             "exchange (~1-2 ms); every GPU is power-capped and the step follows the slowest one"}
        del mc_s, gath_s

    # ---- the same step with every projection as the reference orders it (no weight folding)
    unfolded = None
    if world == 1 and args.unfolded_steps > 0 and fold:
        u_ms = timed(make_step(cb, plans, None, None, 0, False), args.unfolded_steps, warm=1)
        unfolded = {"ms_per_step": u_ms, "value": S * vsec(w) / (u_ms * 1e-3), "unit": "video-s/s", "steps": args.unfolded_steps,
                    "desc": "fold = 0: mm_projector.2 and audio_proj run on every frame, K/V from the d_llm-wide tokens"}

    # ---- end to end: the towers' outputs in pinned host memory, compressed tokens back in host memory
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((rows, K, d), dtype=torch.bfloat16, pin_memory=pinned)
        kw = dict(input_ids=None if ids_dev is None else ids_dev.cpu(), num_query=K, fold=fold, static_out=static_out,
                  chunks_per_batch=min(args.e2e_chunks_per_batch, max(1, S // 6)),
                  taper_head=not args.e2e_no_head_taper)
        if world > 1:
            kw["out_device"] = torch.empty((rows, K, d), dtype=torch.bfloat16, device=dev)
            if gathered is None:
                gathered = torch.empty((world, rows, K, d), dtype=torch.bfloat16, device=dev)

        def e2e_step():
            eng.compress_frames_host(frames_host, audio_host, chunk_start, chunk_len, out_host, **kw)
            if world > 1:   # (the result is read back to the host per range anyway: plain NCCL exchange here)
                dist.all_gather_into_tensor(gathered.view(world * rows, K, d), kw["out_device"])
        e2e_ms = timed(e2e_step, args.e2e_steps, warm=1)
        # the streamed result equals the resident one
        same = bool(torch.equal(out_host[:64].to(dev), compute_range(plan_range(0, 64 // (F_ - 1) + 1), 0,
                                                                     64 // (F_ - 1) + 1, fold, want_static=False)[:64]))
        h2d = frames_host.numel() * 2 + (audio_host.numel() * 2 if Ta else 0) + S * 4 + 2 * rows * 4
        e2e = {"value": world * S * vsec(w) / (e2e_ms * 1e-3), "unit": "video-s/s", "ms_per_step": e2e_ms, "steps": args.e2e_steps,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_host.numel() * 2,
               "host_memory": "pinned" if pinned else "pageable", "matches_resident": same,
               "api": "QFormerEngine.compress_frames_host (tdc_compress_frames per range of chunks; H2D of the "
                      "towers' outputs / compute / D2H of the compressed tokens on 3 streams)"}

    if rank != 0:
        return None

    peaks = measured_peaks()
    model_s, exec_s, kv_exec_s = frames_flops(w, 144, d_in, 768)
    if not fold:
        f_row, f_row_kv = flops_per_row(L, d, K, T, d)
        exec_s = model_s
        kv_exec_s = (F_ - 1) * 2 * (144 + Ta) * d * 2 * H * N_CROSS
    kv_ms, kv_n = prof["kv_gemm"]["ms"], prof["kv_gemm"]["launches"]
    kv_flops_total = kv_exec_s * S * psteps
    kv_achieved = kv_flops_total / (kv_ms * 1e-3) / 1e12 if kv_ms > 0 else None
    line = {
        "metric": "video_seconds_per_sec", "value": value, "unit": "video-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["label"], "entry": "frames (tdc_compress_frames)",
                   "boundary": "tower features [frames, 144, 1024] + audio tokens [frames, 50, 768] -> mm_projector, "
                               "image_newline, audio_proj, query build, Q-Former, vision_proj + L2-normalise -> "
                               "key-frame tokens [chunks, 206, d] + compressed tokens [rows, K, d]",
                   "segments_per_gpu": S, "frames_per_gpu": n_frames, "rows_per_gpu": rows, "kv_tokens": L, "d_enc": d,
                   "d_out": d, "d_frame_in": d_in, "num_query": K, "num_text": T, "fold": fold,
                   "parallelism": f"dp{world} (video-second ranges per GPU)", "exchange": exchange,
                   "l2": f"inputs {(frames_dev.numel() + (audio_dev.numel() if Ta else 0)) * 2 / 1e9:.1f} GB per GPU "
                         f">> 126 MB L2 (no flush needed)",
                   "accumulate": "fp32 (TMEM), LN/softmax/residual fp32"},
        "clocks": clocks, "exchange_check": exchange_check, "strong": strong, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": {"kernel": "tdc_gemm_kernel (cross-attn K/V projection of all 6 layers, N=9216: visual tokens "
                               "K=3584 + audio tokens K=768 per row batch)",
                     "bound": "tensor", "achieved": kv_achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": (kv_achieved / peaks["tflops_sustained"]) if kv_achieved else None,
                     "traffic": (measured_traffic_per_row("kv_gemm_traffic_r02.json") * rows * psteps / max(kv_n // 2, 1))
                     if (fold and measured_traffic_per_row("kv_gemm_traffic_r02.json")) else None,
                     "traffic_note": "DRAM read + write of ONE visual-token K/V launch = ncu bytes per row "
                                     "(profiles/kv_gemm_traffic_r02.json) x rows per launch; the audio-token launches "
                                     "(K = 768, ~8 % of the class time) are not in this figure; algorithmic = "
                                     "rows*144*(3584 + 9216)*2 B + the 66 MB weight",
                     "algorithmic_bytes": rows * psteps / max(kv_n // 2, 1) * 144 * (d + 2 * H * N_CROSS) * 2 + 2 * H * N_CROSS * d * 2,
                     "flops": "EXECUTED by these launches (folded weights): rows x (2*144*3584 + 2*50*768) x 9216",
                     "peak_source": peaks["source"] + " bf16_tflops_sustained", "launches": kv_n,
                     "avg_launch_ms": kv_ms / max(kv_n, 1), "share_of_step": kv_ms / psteps / ms_step},
        "path": {"model_tflop_per_step": model_s * S / 1e12, "executed_tflop_per_step": exec_s * S / 1e12,
                 "model_tflops": model_s * S / (ms_step * 1e-3) / 1e12,
                 "executed_tflops": exec_s * S / (ms_step * 1e-3) / 1e12,
                 "executed_frac_of_sustained_peak": exec_s * S / (ms_step * 1e-3) / 1e12 / peaks["tflops_sustained"],
                 "model_frac_of_sustained_peak": model_s * S / (ms_step * 1e-3) / 1e12 / peaks["tflops_sustained"],
                 "gflop_per_video_second": {"model": model_s / vsec(w) / 1e9, "executed": exec_s / vsec(w) / 1e9},
                 "kernel_ms_per_step": {k: v["ms"] / psteps for k, v in prof.items()}},
        "ranks": per_rank, "unfolded": unfolded,
    }
    if args.parity_rows > 0:
        # the timed workload's own first chunks against the CPU oracle (checker only)
        from oracle import frames_oracle, qformer_oracle as oracle
        n_c = max(1, args.parity_rows // (F_ - 1))
        got = compute_range(plan_range(0, n_c), 0, n_c, fold, want_static=False).float().cpu()
        sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
        ref_st, ref = frames_oracle.frames_stage(sd_t, geom, frames_host[:n_c * F_].float(),
                                                 None if not Ta else audio_host[:n_c * F_].float(), chunk_start[:n_c],
                                                 chunk_len[:n_c], K, None if ids_dev is None else ids_dev.cpu())
        pm = oracle.parity_metrics(got, ref)
        ps = oracle.parity_metrics(static_out[:n_c].float().cpu(), ref_st)
        okf = lambda m_: bool(m_["min_cos"] >= 0.999 and m_["max_abs_over_max_ref"] <= 2e-2 and m_["max_tok_rel_l2"] <= 2e-2)
        line["parity_sample"] = dict(chunks=n_c, rows=int(got.shape[0]), num_query=K, **pm, ok=okf(pm),
                                     key_frames=dict(**ps, ok=okf(ps)))
    if not args.no_cpu_baseline:
        passes = args.cpu_passes or 12
        v, dt, cores = cpu_baseline(geom, sd, w, args.cpu_sample_rows, 99, passes)
        line["cpu_baseline"] = {"value": v, "unit": "video-s/s", "cores": cores, "kind": "port",
                                "sample": f"{passes} x {args.cpu_sample_rows // (F_ - 1) * vsec(w)} video-seconds of the same "
                                          f"workload in {dt:.1f} s (oracle port of the reference from the towers' "
                                          f"outputs, fp32 torch, one batched call per pass)"}
    return line


def bench_tokens(args, w, ctx):
    """Q-Former + vision_proj + L2-normalise on already projected frame tokens (tdc_compress): round 1's line and
    the variant workloads.  Returns the JSON line (rank 0; None elsewhere)."""
    from tdc_video_b200 import QFormerEngine
    world, rank, local_rank, dev, dist = ctx["world"], ctx["rank"], ctx["local_rank"], ctx["dev"], ctx["dist"]
    geom, sd, rows = build_problem(w, 1234)
    L, K, d_enc, d_out = w["kv_tokens"], w["num_query"], w["d_enc"], w["d_out"]
    S = w["segments"]
    T = w["num_text"]
    eng = QFormerEngine(d_enc=d_enc, d_out=d_out, vocab=30522 if T else 0, device=dev, gemm_cta_group=args.cta_group)
    eng.load_weights({k: v for k, v in sd.items() if not k.startswith("mm_projector.")})
    mlp = None
    if w["projector"] == "gelu_mlp":
        from tdc_video_b200.projector import gelu_mlp
        mlp = [torch.from_numpy(sd[f"mm_projector.{i}.{p}"]).to(dev, torch.bfloat16 if p == "weight" else torch.float32)
               for i in (0, 2) for p in ("weight", "bias")]
        args.no_e2e = True          # compress_host streams the vision_proj flavour only
        args.no_multicast = True

    def compute(enc_rows, qs_rows, ts_rows):
        """One pass of the hot path over a row range: Q-Former + projector flavour of the workload."""
        if mlp is None:
            return eng.compress(q_dev, enc_rows, ids_dev, query_set=qs_rows, text_set=ts_rows, out_dtype=torch.bfloat16)
        hidden = eng.forward(q_dev, enc_rows, ids_dev, query_set=qs_rows, text_set=ts_rows, out_dtype=torch.bfloat16)
        return gelu_mlp(hidden[:, :K], *mlp)

    # ---- synthetic inputs, generated on the host in pinned memory (also the e2e source), then made resident
    # (values are drawn with the device RNG for speed — 8e9 normals — then the HOST copy is the
    #  source of truth: the resident tensor is uploaded from it, and e2e streams it every step)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    pinned = not args.no_e2e
    enc_host = None
    if not args.no_e2e:
        try:
            enc_host = torch.empty((rows, L, d_enc), dtype=torch.bfloat16, pin_memory=True)
        except RuntimeError as e:  # e.g. cudaHostAlloc limit on a box with many ranks: pageable staging instead
            print(f"[bench] rank {rank}: pinned allocation failed ({e}); using pageable host memory", file=sys.stderr)
            pinned = False
            enc_host = torch.empty((rows, L, d_enc), dtype=torch.bfloat16)
    enc = torch.empty((rows, L, d_enc), dtype=torch.bfloat16, device=dev)
    chunk = 1024
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        blk = torch.randn((r1 - r0, L, d_enc), generator=g, device=dev)
        blk[:, L - w["audio_tokens"]:] *= 0.5
        enc[r0:r1].copy_(blk.to(torch.bfloat16))
        if enc_host is not None:
            enc_host[r0:r1].copy_(enc[r0:r1])
    del blk
    q_sets = torch.randn((S, K, H), generator=g, device=dev).cpu()    # one query set per segment
    query_set = (torch.arange(rows) // (w["frames_per_segment"] - 1)).to(torch.int32)
    q_dev, qs_dev = q_sets.to(dev), query_set.to(dev)
    ids_dev = ts_dev = None
    if T > 0:   # one BERT-tokenised prompt per video, shared by every row (cambrian_arch.py:1532, 1643-1644)
        ids_dev = torch.randint(1000, 30000, (1, T), generator=torch.Generator().manual_seed(7)).to(dev)
        ts_dev = torch.zeros(rows, dtype=torch.int32, device=dev)
    gathered = torch.empty((world, rows, K, d_out), dtype=torch.bfloat16, device=dev) if world > 1 else None
    nb = max(1, args.gather_batches)
    bounds = [(rows * b // nb, rows * (b + 1) // nb) for b in range(nb)]
    mcast, exchange = None, "none (1 GPU)"
    if world > 1:
        exchange = "NCCL all-gather of compressed tokens"
        if not args.no_multicast:
            try:
                from tdc_video_b200.dist import MulticastGather
                mcast = MulticastGather(rows, (K, d_out), torch.bfloat16, dev)
                exchange = "NVSwitch multicast stores (multimem.st) from the final kernel + device barrier"
            except Exception as e:  # transport fallback only; compute path is identical
                if rank == 0:
                    print(f"[bench] symmetric-memory multicast unavailable ({type(e).__name__}: {e}); using NCCL",
                          file=sys.stderr)
                mcast = None

    def step():
        if world == 1:
            return compute(enc, qs_dev, ts_dev)
        # the path's one exchange step: all-gather of the compressed tokens, issued per row batch on
        # NCCL's stream so that it overlaps the next batch's kernels; every rank ends with the
        # rank-ordered sequence [world, rows, K, d_out]
        if mcast is not None:
            # all-gather fused into the producing kernel: the L2-normalise kernel stores every row through
            # the multicast mapping, so all ranks receive it while the kernel runs
            for r0, r1 in bounds:
                eng.compress_multicast(q_dev, enc[r0:r1], mcast.slot_ptr(r0), ids_dev, query_set=qs_dev[r0:r1],
                                       text_set=None if ts_dev is None else ts_dev[r0:r1], out_dtype=torch.bfloat16)
            mcast.barrier()
            return mcast.gathered
        works = []
        for r0, r1 in bounds:
            out = compute(enc[r0:r1], qs_dev[r0:r1], None if ts_dev is None else ts_dev[r0:r1])
            works.append(dist.all_gather([gathered[w, r0:r1] for w in range(world)], out, async_op=True))
        for wk in works:
            wk.wait()
        return gathered.view(world * rows, K, d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        out = step()
    barrier()
    assert torch.isfinite(out[:8].float()).all()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0   # our kernels only (NCCL's are not counted)
    # per-kernel-class device times come from a SEPARATE profiled pass (two CUDA events around every launch would
    # otherwise sit inside the timed region: ~1800 event records per step)
    psteps = max(1, min(args.steps, 5))
    eng.set_profiling(True)
    eng.reset_profile()
    for _ in range(psteps):
        step()
    barrier()
    prof = eng.profile()
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    # every rank's own device time and kernel-time sum (the step ends at a barrier, so `value` follows the slowest GPU)
    mine = torch.tensor([ms / args.steps, sum(v["ms"] for v in prof.values()) / psteps], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine)
    per_rank = {"ms_per_step": [round(float(x[0]), 2) for x in per_rank],
                "kernel_ms_per_step": [round(float(x[1]), 2) for x in per_rank]}
    ms_step = float(t.item()) / args.steps
    value = world * S * vsec(w) / (ms_step * 1e-3)

    # ---- the exchange verifies itself: every rank's locally computed rows, all-gathered by NCCL, must equal
    # bit for bit what the timed step left in the gather buffer on EVERY rank (multicast stores + device barrier,
    # or the per-batch NCCL all-gathers) — rows are independent, so batching cannot change a bit
    exchange_check = None
    strong = None
    if world > 1:
        final = step()
        barrier()
        local_all = compute(enc, qs_dev, ts_dev)
        ref_gather = torch.empty((world * rows, K, d_out), dtype=torch.bfloat16, device=dev)
        dist.all_gather_into_tensor(ref_gather, local_all.contiguous())
        ok = torch.equal(final.reshape(world * rows, K, d_out), ref_gather)
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange_check = bool(flag.item())
        del ref_gather, local_all, final

        # ---- strong scaling (BASELINE config 4): ONE video of S segments sharded over the N GPUs, S/N contiguous
        # video-seconds per GPU, the same fused exchange; reference point = one GPU compressing all S segments
        # (no exchange), timed in the same run on every rank (max over ranks, like everything else)
        rows_s = rows // world
        sb = [(rows_s * b // nb, rows_s * (b + 1) // nb) for b in range(nb)]
        mc_s, gath_s = None, None
        if mcast is not None:
            from tdc_video_b200.dist import MulticastGather
            mc_s = MulticastGather(rows_s, (K, d_out), torch.bfloat16, dev)
        else:
            gath_s = torch.empty((world, rows_s, K, d_out), dtype=torch.bfloat16, device=dev)
        lo = rank * rows_s                         # this rank's range of the video

        def strong_step():
            if mc_s is not None:
                for r0, r1 in sb:
                    eng.compress_multicast(q_dev, enc[lo + r0:lo + r1], mc_s.slot_ptr(r0), ids_dev,
                                           query_set=qs_dev[lo + r0:lo + r1],
                                           text_set=None if ts_dev is None else ts_dev[lo + r0:lo + r1],
                                           out_dtype=torch.bfloat16)
                mc_s.barrier()
                return mc_s.gathered
            works = []
            for r0, r1 in sb:
                o = compute(enc[lo + r0:lo + r1], qs_dev[lo + r0:lo + r1],
                            None if ts_dev is None else ts_dev[lo + r0:lo + r1])
                works.append(dist.all_gather([gath_s[w_, r0:r1] for w_ in range(world)], o, async_op=True))
            for wk in works:
                wk.wait()
            return gath_s.view(world * rows_s, K, d_out)

        def timed(fn, n):
            for _ in range(2):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            barrier()
            tt = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        strong_ms = timed(strong_step, args.strong_steps)
        n1_ms = timed(lambda: compute(enc, qs_dev, ts_dev), max(2, args.strong_steps // 2))
        # the sharded video equals the same rows compressed by one GPU (rank 0's copy of the check: its own range)
        mine_s = compute(enc[lo:lo + rows_s], qs_dev[lo:lo + rows_s], None if ts_dev is None else ts_dev[lo:lo + rows_s])
        got = strong_step().reshape(world, rows_s, K, d_out)[rank]
        flag = torch.tensor([int(torch.equal(got, mine_s))], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        strong = {"segments_total": S, "segments_per_gpu": S // world, "rows_per_gpu": rows_s, "ms_per_step": strong_ms,
                  "value": S * vsec(w) / (strong_ms * 1e-3), "unit": "video-s/s", "one_gpu_ms_per_step": n1_ms,
                  "speedup_vs_n1": n1_ms / strong_ms, "steps": args.strong_steps, "exchange": exchange,
                  "own_rows_match": bool(flag.item()),
                  "limiter": "per-GPU step = compute of S/N segments + device barrier closing the multicast exchange; "
                             "every GPU is power-capped and the step follows the slowest one"}
        del mc_s, gath_s

    # ---- end to end: KV tokens in pinned host memory, result back in host memory
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((rows, K, d_out), dtype=torch.bfloat16, pin_memory=pinned)
        e2e_kw = dict(query_set=query_set, rows_per_batch=args.e2e_rows_per_batch,
                      input_ids=None if ids_dev is None else ids_dev.cpu(),
                      text_set=None if ts_dev is None else ts_dev.cpu())
        if world > 1:   # the rank's own rows stay on the GPU too: send buffer of the exchange
            e2e_kw["out_device"] = torch.empty((rows, K, d_out), dtype=torch.bfloat16, device=dev)
        eng.compress_host(q_sets, enc_host, out_host, **e2e_kw)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.e2e_steps):
            eng.compress_host(q_sets, enc_host, out_host, **e2e_kw)
            if world > 1:
                dist.all_gather_into_tensor(gathered.view(world * rows, K, d_out), e2e_kw["out_device"])
                # (e2e keeps the plain NCCL exchange: the result is read back to the host per batch anyway)
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te.item()) / args.e2e_steps
        e2e = {"value": world * S * vsec(w) / (e2e_ms * 1e-3), "unit": "video-s/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": enc_host.numel() * 2 + q_sets.numel() * 4 + query_set.numel() * 4,
               "d2h_bytes_per_step": out_host.numel() * 2, "host_memory": "pinned" if pinned else "pageable",
               "api": "QFormerEngine.compress_host (tdc_compress per row batch, H2D / compute / D2H on 3 streams)"}

    if rank != 0:
        return None

    peaks = measured_peaks()
    f_row, f_row_kv = flops_per_row(L, d_enc, K, T, d_out, w["projector"])
    kv_ms, kv_n = prof["kv_gemm"]["ms"], prof["kv_gemm"]["launches"]
    kv_flops_per_launch = f_row_kv * rows * psteps / max(kv_n, 1)
    kv_achieved = kv_flops_per_launch / (kv_ms / max(kv_n, 1) * 1e-3) / 1e12 if kv_ms > 0 else None
    path_tflops = f_row * rows / (ms_step * 1e-3) / 1e12
    line = {
        "metric": "video_seconds_per_sec", "value": value, "unit": "video-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "desc": w["label"], "segments_per_gpu": S, "rows_per_gpu": rows,
                   "kv_tokens": L, "d_enc": d_enc, "d_out": d_out, "num_query": K, "num_text": T, "projector": w["projector"],
                   "parallelism": f"dp{world} (video-second ranges per GPU)", "exchange": exchange,
                   "l2": f"inputs {enc.numel() * 2 / 1e9:.1f} GB per GPU >> 126 MB L2 (no flush needed)",
                   "accumulate": "fp32 (TMEM), LN/softmax/residual fp32"},
        "clocks": clocks,
        "exchange_check": exchange_check,
        "strong": strong,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"kernel": "tdc_gemm_kernel (cross-attn K/V projection, all 6 layers, N=9216)", "bound": "tensor",
                     "achieved": kv_achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": (kv_achieved / peaks["tflops_sustained"]) if kv_achieved else None,
                     "traffic": (measured_traffic_per_row() * rows * psteps / max(kv_n, 1))
                     if measured_traffic_per_row() else None,
                     "traffic_note": "bytes per launch = ncu dram read+write per row (profiles/kv_gemm_traffic.json) x rows "
                                     "per launch; algorithmic = rows*L*(d_enc + 9216)*2 B",
                     "algorithmic_bytes": rows * psteps / max(kv_n, 1) * L * (d_enc + 2 * H * N_CROSS) * 2,
                     "peak_source": peaks["source"] + " bf16_tflops_sustained", "launches": kv_n,
                     "avg_launch_ms": kv_ms / max(kv_n, 1), "share_of_step": kv_ms / psteps / ms_step},
        "path": {"algorithmic_tflops": path_tflops, "frac_of_sustained_peak": path_tflops / peaks["tflops_sustained"],
                 "gflop_per_row": f_row / 1e9,
                 "kernel_ms_per_step": {k: v["ms"] / psteps for k, v in prof.items()}},
        "ranks": per_rank,
    }
    if args.parity_rows > 0 and mlp is None:
        # the timed workload's own rows against the CPU oracle (checker only): the first rows of this rank's video
        from oracle import qformer_oracle as oracle
        n_chk = min(args.parity_rows, rows)
        got = compute(enc[:n_chk], qs_dev[:n_chk], None if ts_dev is None else ts_dev[:n_chk]).float().cpu()
        sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
        ids_chk = None if ids_dev is None else ids_dev.cpu().expand(n_chk, -1)
        ref = oracle.compress(sd_t, geom, q_sets[query_set[:n_chk].long()], enc[:n_chk].float().cpu(), ids_chk)
        pm = oracle.parity_metrics(got, ref)
        line["parity_sample"] = dict(rows=n_chk, num_query=K, **pm,
                                     ok=bool(pm["min_cos"] >= 0.999 and pm["max_abs_over_max_ref"] <= 2e-2
                                             and pm["max_tok_rel_l2"] <= 2e-2))
    if not args.no_cpu_baseline:
        passes = args.cpu_passes or 16
        v, dt, cores = cpu_baseline(geom, sd, w, args.cpu_sample_rows, 99, passes)
        line["cpu_baseline"] = {"value": v, "unit": "video-s/s", "cores": cores, "kind": "port",
                                "sample": f"{passes} x {args.cpu_sample_rows} rows of the same workload in {dt:.1f} s "
                                          f"(oracle port of the reference, fp32 torch, {args.cpu_sample_rows}-row batches)"}
        if w["projector"] == "vision_proj":
            line["cpu_baseline"]["reference_batching"] = {
                "rows_per_call": 7, "unit": "video-s/s",
                "value": cpu_baseline_reference_batching(geom, sd, w, args.cpu_sample_rows, 99)}
    return line


if __name__ == "__main__":
    main()
