"""torchrun --nproc-per-node N tools/dist_check.py — sharded TDCCompressor.compress_video on N GPUs must
equal the single-GPU result bit for bit (rows are independent)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200.compressor import TDCCompressor  # noqa: E402
from tdc_video_b200.qformer import QFormerConfig  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)  # identical module + inputs on every rank
    cfg = QFormerConfig(vocab_size=64, hidden_size=128, num_hidden_layers=4, num_attention_heads=2,
                        intermediate_size=256, max_position_embeddings=16)
    comp = TDCCompressor(96, context_token_num=16, audio_input=True, qformer_config=cfg).cuda().eval()
    sizes = [5, 1, 19, 8, 2, 11, 30, 3]
    n = sum(sizes)
    frames = torch.randn(n, 30, 96, device="cuda", dtype=torch.bfloat16)
    audio = torch.randn(n, 6, 768, device="cuda", dtype=torch.bfloat16)
    ids = torch.randint(0, 64, (1, 5), device="cuda")
    single = comp.compress_video(frames, sizes, input_ids=ids, audio_frames=audio)
    sharded = comp.compress_video(frames, sizes, input_ids=ids, audio_frames=audio, shard=True)
    torch.cuda.synchronize()
    ok = torch.equal(single, sharded)
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_check world={world}: sharded == single: {bool(flag.item())} ({tuple(single.shape)})")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
