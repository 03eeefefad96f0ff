"""Latency of the TDC stage for ONE reference-sized video (what main.py / eval_*.py do per question):
224 frames (the reference's cap, cambrian_arch.py:908), 24 adaptive boundaries, Qwen2-7B widths, 156 visual + 50
audio tokens per frame, K = 16, a 32-token prompt -> `pipeline.tdc_video_stage` (adapt_segment, audio pooling,
chunked Q-Former compression, assembly).  Wall clock around the call incl. all host logic, after warm-up.
Prints one JSON line; run on a B200: python tools/bench_latency.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200.compressor import TDCCompressor  # noqa: E402
from tdc_video_b200.pipeline import tdc_video_stage  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    n, d, T = 224, 3584, 32
    comp = TDCCompressor(d, context_token_num=16, query_type="Avg_pool", text_input=True, add_static=True,
                         audio_input=True).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(5)
    g0 = torch.Generator(device=dev).manual_seed(6)
    frames = torch.randn(n, 156, d, device=dev, generator=g).to(torch.bfloat16)
    # DINO features: slow drift + a jump every ~9 frames so that the 24 boundaries are well defined
    drift = torch.cumsum(torch.randn(n, 1, 1536, device=dev, generator=g) * 0.05, 0)
    dino = (torch.randn(1, 576, 1536, device=dev, generator=g) + drift)
    dino[::9] += torch.randn(len(range(0, n, 9)), 1, 1536, device=dev, generator=g)
    dino = dino.to(torch.bfloat16)
    windows = [torch.randn(1, 500 if w < 22 else 200, 768, device=dev, generator=g).to(torch.bfloat16) for w in range(23)]
    flags = [1] * n
    ids = torch.randint(1000, 30000, (1, T), device=dev, generator=g)
    kw = dict(input_ids=ids, audio_windows=windows, sample_indices=flags, max_visual_len=32000)

    def run():
        return tdc_video_stage(comp, frames, dino, **kw)

    for _ in range(3):
        out = run()
    torch.cuda.synchronize()
    walls, devs = [], []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        out = run()
        e1.record()
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - t0) * 1e3)
        devs.append(e0.elapsed_time(e1))
    _, comp_rows, plan = comp.compress_video(frames, [9] * 24 + [8], input_ids=ids, return_parts=True)

    # the Q-Former call alone (tdc_compress on the same rows): eager launches vs one CUDA-graph replay
    prep = comp._prepare(frames, [9] * 24 + [8], ids, None)
    eng = comp._engine()
    R = prep["plan"].num_rows
    args = (prep["q_sets"], prep["enc"], prep["ids"])
    kws = dict(query_set=prep["query_set"].cuda(), text_set=torch.zeros(R, dtype=torch.int32, device=dev),
               out_dtype=torch.bfloat16)

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

    eager_dev, eager_wall = timed(lambda: eng.compress(*args, **kws))
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            eng.compress(*args, **kws)
    torch.cuda.current_stream().wait_stream(side)
    graph_dev, graph_wall = timed(g.replay)
    engine = {"rows": int(R), "eager_ms": eager_dev, "eager_wall_ms": eager_wall, "graph_ms": graph_dev,
              "graph_wall_ms": graph_wall}

    # the same video from the TOWERS' outputs (round 2): mm_projector + newline + audio_proj + queries + Q-Former in one
    # tdc_compress_frames call (weights folded), adapt_segment and audio pooling as above
    comp2 = TDCCompressor(d, context_token_num=16, query_type="Avg_pool", text_input=True, add_static=True,
                          audio_input=True, mm_input_size=1024).to(dev).eval()
    feats = torch.randn(n, 144, 1024, device=dev, generator=g0).to(torch.bfloat16)

    def run2():
        return tdc_video_stage(comp2, None, dino, tower_features=feats, **kw)

    for _ in range(3):
        out2 = run2()
    torch.cuda.synchronize()
    walls2, devs2 = [], []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        out2 = run2()
        e1.record()
        torch.cuda.synchronize()
        walls2.append((time.perf_counter() - t0) * 1e3)
        devs2.append(e0.elapsed_time(e1))
    # tdc_compress_frames alone on that video's plan: eager launches vs one CUDA-graph replay (the call is
    # stream-ordered and makes no host reads, so it captures as is)
    from tdc_video_b200.compressor import plan_chunks
    pl = plan_chunks([9] * 24 + [8], True)
    i32 = lambda a: torch.from_numpy(a.astype(np.int32)).to(dev)
    sf, rf, rc = i32(pl.static_frames), i32(pl.row_frames), i32(pl.row_chunk)
    aud = (torch.randn(n, 50, 768, device=dev, generator=g0) * 0.5).to(torch.bfloat16)
    eng2 = comp2._frames_engine()
    call = lambda: eng2.compress_frames(feats, sf, rf, rc, audio=aud, input_ids=ids, num_query=16)
    f_eager_dev, f_eager_wall = timed(call)
    l0 = eng2.launch_count()
    ref_static, ref_out = call()
    launches = eng2.launch_count() - l0
    g2 = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g2, stream=side):
            cap_static, cap_out = call()
    torch.cuda.current_stream().wait_stream(side)
    f_graph_dev, f_graph_wall = timed(g2.replay)
    torch.cuda.synchronize()
    frames_call = {"rows": int(pl.num_rows), "chunks": int(pl.num_chunks), "kernel_launches": int(launches),
                   "eager_ms": f_eager_dev, "eager_wall_ms": f_eager_wall, "graph_ms": f_graph_dev,
                   "graph_wall_ms": f_graph_wall,
                   "graph_equals_eager": bool(torch.equal(cap_out, ref_out) and torch.equal(cap_static, ref_static))}
    from_towers = {"wall_ms_median": float(np.median(walls2)), "device_ms_median": float(np.median(devs2)),
                   "output_tokens": int(out2.shape[0]), "finite": bool(torch.isfinite(out2.float()).all()),
                   "includes": "mm_projector on 224 frames (1.07 TFLOP), newline, audio_proj, query build",
                   "tdc_compress_frames_alone": frames_call}
    print(json.dumps({
        "metric": "TDC stage latency, one 224-frame video (Qwen2-7B widths, L=206, K=16, T=32, audio)",
        "wall_ms_median": float(np.median(walls)), "wall_ms_min": float(min(walls)),
        "device_ms_median": float(np.median(devs)), "output_tokens": int(out.shape[0]),
        "rows": int(plan.num_rows), "chunks": int(plan.num_chunks),
        "reference_launch_pattern": f"{plan.num_chunks} Q-Former calls of <= 7 rows (cambrian_arch.py:1603-1692)",
        "tdc_compress_alone": engine, "from_towers": from_towers, "finite": bool(torch.isfinite(out.float()).all())}))


if __name__ == "__main__":
    main()
