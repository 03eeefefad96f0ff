"""torchrun probe: does torch symmetric memory (peer pointers / NVSwitch multicast) work on this box,
and how fast are NCCL all-gather vs direct peer copies for the bench payload?"""
import os

import torch
import torch.distributed as dist


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 10800 * 16 * 3584
    src = torch.randn(n, device=dev).bfloat16()
    dst = torch.empty(world * n, dtype=torch.bfloat16, device=dev)
    ms = timed(lambda: dist.all_gather_into_tensor(dst, src))
    if rank == 0:
        print(f"NCCL all_gather_into_tensor {n * 2 / 1e9:.2f} GB/rank, world {world}: {ms:.2f} ms "
              f"({(world - 1) * n * 2 / ms / 1e6:.0f} GB/s in per rank)", flush=True)
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(world * n, dtype=torch.bfloat16, device=dev)
        h = symm.rendezvous(t, dist.group.WORLD.group_name)
        if rank == 0:
            print("symm_mem ok: world", h.world_size, "multicast", h.has_multicast_support, hex(h.multicast_ptr or 0),
                  "ptrs", [hex(p) for p in h.buffer_ptrs], flush=True)

        def push():
            for p in range(world):
                peer = h.get_buffer(p, (world * n,), torch.bfloat16)
                peer[rank * n:(rank + 1) * n].copy_(src)
            h.barrier()

        ms2 = timed(push)
        if rank == 0:
            print(f"symm_mem peer copies (copy kernels/CE) + barrier: {ms2:.2f} ms "
                  f"({(world - 1) * n * 2 / ms2 / 1e6:.0f} GB/s out per rank)", flush=True)
        ok = torch.equal(t[rank * n:(rank + 1) * n], src)
        other = (rank + 1) % world
        # what the neighbour wrote into my buffer must equal what it holds: compare via checksum all-reduce
        s_local = src.float().sum()
        sums = [torch.zeros_like(s_local) for _ in range(world)]
        dist.all_gather(sums, s_local)
        ok2 = torch.allclose(t[other * n:(other + 1) * n].float().sum(), sums[other], rtol=1e-3)
        if rank == 0:
            print("symm_mem data check:", ok, ok2, flush=True)
    except Exception as e:  # noqa
        if rank == 0:
            print("symm_mem FAILED:", type(e).__name__, str(e)[:400], flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
