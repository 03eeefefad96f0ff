#!/usr/bin/env python
"""bench.py — TDC compression throughput (video-seconds/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hour_qwen7b|cfg2_llama3b|...] [--num-text T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

One "step" = one pass of the hot path over one synthetic video per GPU: for every dynamic
frame (row) the Q-Former (12 layers, cross-attention to the frame's L KV tokens) + vision_proj
+ L2-normalise, i.e. tdc/cambrian_arch.py:1603-1692 for all chunks at once, then (N > 1) the
all-gather of the compressed tokens so that every rank holds the ordered sequence (fused into the
final kernel through NVSwitch multicast stores; `--no-multicast` = NCCL all-gather).

 * `value`  : whole-job video-seconds/s, inputs resident in HBM, CUDA-event timed, max over ranks
 * `e2e`    : the same through `QFormerEngine.compress_host` with the KV tokens in pinned HOST
              memory (H2D of every row inside the timed region, result copied back to host)
 * `roofline`: dominant kernel = the cross-attention K/V projection GEMM (tcgen05), algorithmic
              FLOPs / its CUDA-event time (events on the launching stream, recorded by the library)
 * `cpu_baseline`: the reference algorithm (oracle port, fp32 torch) on the host cores, bounded sample

Synthetic data, random-init weights (tdc_video_b200/synth.py statistics); weak scaling: every GPU
compresses its own `segments` video-seconds.
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


def measured_traffic_per_row():
    """DRAM bytes per row of the KV-projection GEMM from the committed ncu --set full capture
    (profiles/kv_gemm_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum over the rows of that launch)."""
    p = os.path.join(ROOT, "profiles", "kv_gemm_traffic.json")
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
    rows = w["segments"] * (w["frames_per_segment"] - 1)
    return geom, sd, rows


def cpu_baseline(geom, sd, w, sample_rows, seed, passes=1):
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
    return rows_per_s / (w["frames_per_segment"] - 1), dt, cores


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
    return sample_rows / dt / (w["frames_per_segment"] - 1)


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
        "config": {"workload": args.workload, "desc": w["label"], "rows_per_step_sample": sample * passes},
        "cpu_baseline": {"value": value, "unit": "video-s/s", "cores": cores, "kind": "port",
                         "sample": f"{passes} x {sample} rows (= {passes * sample / (w['frames_per_segment'] - 1):.1f} "
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
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.segments:
        w["segments"] = args.segments
    w["num_text"] = args.num_text
    if args.num_query:
        w["num_query"] = args.num_query
    w.setdefault("projector", "vision_proj")
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
    try:
        enc_host = torch.empty((rows, L, d_enc), dtype=torch.bfloat16, pin_memory=pinned)
    except RuntimeError as e:  # e.g. cudaHostAlloc limit on a box with many ranks: pageable staging instead
        print(f"[bench] rank {rank}: pinned allocation failed ({e}); using pageable host memory", file=sys.stderr)
        pinned = False
        enc_host = torch.empty((rows, L, d_enc), dtype=torch.bfloat16)
    chunk = 1024
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        blk = torch.randn((r1 - r0, L, d_enc), generator=g, device=dev)
        blk[:, L - w["audio_tokens"]:] *= 0.5
        enc_host[r0:r1].copy_(blk.to(torch.bfloat16))
    del blk
    q_sets = torch.randn((S, K, H), generator=g, device=dev).cpu()    # one query set per segment
    query_set = (torch.arange(rows) // (w["frames_per_segment"] - 1)).to(torch.int32)
    enc = enc_host.to(dev)
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
    eng.set_profiling(True)
    eng.reset_profile()
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
    prof = eng.profile()
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    # every rank's own device time and kernel-time sum (the step ends at a barrier, so `value` follows the slowest GPU)
    mine = torch.tensor([ms / args.steps, sum(v["ms"] for v in prof.values()) / args.steps], dtype=torch.float64, device=dev)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine)
    per_rank = {"ms_per_step": [round(float(x[0]), 2) for x in per_rank],
                "kernel_ms_per_step": [round(float(x[1]), 2) for x in per_rank]}
    ms_step = float(t.item()) / args.steps
    value = world * S / (ms_step * 1e-3)

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
                  "value": S / (strong_ms * 1e-3), "unit": "video-s/s", "one_gpu_ms_per_step": n1_ms,
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
        e2e = {"value": world * S / (e2e_ms * 1e-3), "unit": "video-s/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": enc_host.numel() * 2 + q_sets.numel() * 4 + query_set.numel() * 4,
               "d2h_bytes_per_step": out_host.numel() * 2, "host_memory": "pinned" if pinned else "pageable",
               "api": "QFormerEngine.compress_host (tdc_compress per row batch, H2D / compute / D2H on 3 streams)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    f_row, f_row_kv = flops_per_row(L, d_enc, K, T, d_out, w["projector"])
    kv_ms, kv_n = prof["kv_gemm"]["ms"], prof["kv_gemm"]["launches"]
    kv_flops_per_launch = f_row_kv * rows * args.steps / max(kv_n, 1)
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
                     "traffic": (measured_traffic_per_row() * rows * args.steps / max(kv_n, 1))
                     if measured_traffic_per_row() else None,
                     "traffic_note": "bytes per launch = ncu dram read+write per row (profiles/kv_gemm_traffic.json) x rows "
                                     "per launch; algorithmic = rows*L*(d_enc + 9216)*2 B",
                     "algorithmic_bytes": rows * args.steps / max(kv_n, 1) * L * (d_enc + 2 * H * N_CROSS) * 2,
                     "peak_source": peaks["source"] + " bf16_tflops_sustained", "launches": kv_n,
                     "avg_launch_ms": kv_ms / max(kv_n, 1), "share_of_step": kv_ms / (ms_step * args.steps)},
        "path": {"algorithmic_tflops": path_tflops, "frac_of_sustained_peak": path_tflops / peaks["tflops_sustained"],
                 "gflop_per_row": f_row / 1e9,
                 "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()}},
        "ranks": per_rank,
    }
    if args.parity_rows > 0 and mlp is None:
        # the timed workload's own rows against the CPU oracle (checker only): the first rows of this rank's video
        from oracle import qformer_oracle as oracle
        n_chk = min(args.parity_rows, rows)
        got = compute(enc[:n_chk], qs_dev[:n_chk], None if ts_dev is None else ts_dev[:n_chk]).float().cpu()
        sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
        ids_chk = None if ids_dev is None else ids_dev.cpu().expand(n_chk, -1)
        ref = oracle.compress(sd_t, geom, q_sets[query_set[:n_chk].long()], enc_host[:n_chk].float(), ids_chk)
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
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
