"""torchrun check: tdc_compress_multicast delivers every rank's rows to all ranks (== NCCL all-gather), and so do
MulticastGather.put_async's copy-engine (tdc_peer_copy) and multimem (tdc_multicast_copy) modes."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200.synth import QFormerGeometry, make_state_dict  # noqa: E402
from tdc_video_b200 import QFormerEngine  # noqa: E402
from tdc_video_b200.dist import MulticastGather  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    geom = QFormerGeometry(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=96, vocab=0)
    eng = QFormerEngine(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=96, device=dev)
    eng.load_weights(make_state_dict(geom, 3, with_text=False))
    rows, K = 37, 16
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    enc = torch.randn((rows, 29, 64), generator=g, device=dev).bfloat16()
    q = torch.randn((rows, K, 128), generator=g, device=dev)
    local_out = eng.compress(q, enc, out_dtype=torch.bfloat16)
    ref = torch.empty((world * rows, K, 96), dtype=torch.bfloat16, device=dev)
    dist.all_gather_into_tensor(ref, local_out)
    mg = MulticastGather(rows, (K, 96), torch.bfloat16, dev)
    mg.buf.zero_()
    mg.barrier()
    for r0, r1 in ((0, 20), (20, rows)):
        eng.compress_multicast(q[r0:r1], enc[r0:r1], mg.slot_ptr(r0), out_dtype=torch.bfloat16)
    mg.barrier()
    torch.cuda.synchronize()
    ok = torch.equal(mg.gathered, ref)
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"mcast_check world={world}: multicast gather == NCCL all-gather: {bool(flag.item())}")
    # put_async: local rows -> every rank's buffer on side streams (copy engines / multimem copy kernel)
    for mode in ("dma", "multimem"):
        mg.buf.zero_()
        dist.barrier()
        torch.cuda.synchronize()
        mg.put_async(local_out[:20], 0, mode=mode)
        mg.put_async(local_out[20:].contiguous(), 20, mode=mode)
        mg.barrier()
        torch.cuda.synchronize()
        f2 = torch.tensor([int(torch.equal(mg.gathered, ref))], device=dev)
        dist.all_reduce(f2, op=dist.ReduceOp.MIN)
        flag = torch.minimum(flag, f2)
        if rank == 0:
            print(f"mcast_check world={world}: put_async(mode={mode}) == NCCL all-gather: {bool(f2.item())}")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
