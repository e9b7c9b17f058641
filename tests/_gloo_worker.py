"""world_size-2 gloo worker for tests/test_host_cpu.py: frame sharding + the output all-gather."""
import os

import torch
import torch.distributed as dist

from matryodshka_b200.runtime import all_gather_frames, shard_frames


def main():
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    n_frames, H, W = 4, 4, 8
    full = torch.arange(n_frames * H * W * 3, dtype=torch.float32).reshape(n_frames, H, W, 3)
    lo, hi = shard_frames(n_frames, rank, ws)
    local = full[lo:hi].clone() * 1.0
    out = all_gather_frames(local, ws)
    assert torch.equal(out, full), "gathered frames must equal the single-process result bit for bit"
    dist.barrier()
    if rank == 0:
        print("GLOO_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
