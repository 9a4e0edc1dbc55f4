#!/usr/bin/env python
"""Multi-GPU check of the fused histogram + NVLink exchange (b200wm_pattern_hist_publish) against the NCCL all-gather
path, and their cost.   torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/publish_check.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "video-fingerprinting_b200"))

import torch                        # noqa: E402
import torch.distributed as dist    # noqa: E402

from b200wm.vote import SegmentVote  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L, per, frames = 8, 50, 3000
    n_seg = per * world
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    plain = SegmentVote(n_seg, L, dev, owned=(rank * per, per))
    fused = SegmentVote(n_seg, L, dev, owned=(rank * per, per), symmetric=True)
    frame_seg = (torch.arange(frames, device=dev, dtype=torch.int32) // 60 + rank * per).contiguous()
    ok = True
    for it in range(5):
        packed = torch.randint(0, 1 << L, (frames,), device=dev, generator=gen, dtype=torch.int64)
        packed[::3] = 0x65
        plain.reset().add(packed, frame_segment=frame_seg, order_offset=rank * frames).combine()
        fused.reset().add(packed, frame_segment=frame_seg, order_offset=rank * frames).combine()
        a, b = plain.result(), fused.result()
        same = all((x[0] is None and y[0] is None) or (x[0].tolist() == y[0].tolist() and x[1] == y[1] and x[2].tolist() == y[2].tolist()
                                                      and x[3] == y[3]) for x, y in zip(a, b))
        same = same and torch.equal(plain.hist.reshape(-1), fused.hist.reshape(-1)) and torch.equal(plain.first_seen.reshape(-1), fused.first_seen.reshape(-1))
        ok = ok and same
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    def cost(vote, sync_combine):
        for _ in range(5):
            vote.reset().add(packed, frame_segment=frame_seg, order_offset=rank * frames).combine()
        torch.cuda.synchronize()
        dist.barrier()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            vote.reset().add(packed, frame_segment=frame_seg, order_offset=rank * frames).combine()
        z.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(z) / 50], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ms_plain, ms_fused = cost(plain, True), cost(fused, True)
    if rank == 0:
        print(json.dumps({"world": world, "identical_to_nccl_all_gather": bool(flag.item()), "block_bytes": plain._block_len * 4,
                          "reset_hist_allgather_ms": ms_plain, "reset_fused_hist_publish_ms": ms_fused}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
