"""Multi-GPU worker (torchrun): frames sharded over ranks, one NCCL all-gather of the rendered
outputs; the gathered result must equal a single-GPU run of all frames bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matryodshka_b200 import synth  # noqa: E402
from matryodshka_b200.runtime import FrameGather, MSIPipeline, all_gather_frames, shard_frames  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    H, W, P, ngf = 64, 128, 32, 64
    n_frames = 2 * world
    ref, src = synth.ods_pair(n_frames, H, W)
    tp = synth.target_positions(n_frames)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    lo, hi = shard_frames(n_frames, rank, world)
    pipe = MSIPipeline(wts, H, W, P, ngf, batch=hi - lo, device=dev)
    pipe.set_inputs(ref[lo:hi], src[lo:hi], tgt_pos=tp[lo:hi])
    pipe.step()
    rgb = all_gather_frames(pipe.out["rgb"], world)
    rgb8 = all_gather_frames(pipe.out["rgb_u8"], world)
    torch.cuda.synchronize()
    if rank == 0:
        full = MSIPipeline(wts, H, W, P, ngf, batch=n_frames, device=dev, use_graph=False)
        full.set_inputs(ref, src, tgt_pos=tp)
        full.step()
        torch.cuda.synchronize()
        assert torch.equal(rgb, full.out["rgb"]), "N-GPU gathered frames differ from the 1-GPU run"
        assert torch.equal(rgb8, full.out["rgb_u8"])
        print(f"NCCL_OK world={world} frames={n_frames}")
    # fused form: the render kernel stores its frames into every rank's gathered buffer (symmetric
    # memory; multimem.st if the fabric has multicast, and plain peer stores) == the NCCL all-gather
    for mode in ("auto", "peer"):
        try:
            g = FrameGather(hi - lo, H, W, dev, mode=mode)
        except Exception as e:  # no symmetric memory on this box: the NCCL path above is the fallback
            if rank == 0:
                print(f"FUSED_SKIP mode={mode}: {type(e).__name__}: {e}")
            continue
        fused = MSIPipeline(wts, H, W, P, ngf, batch=hi - lo, device=dev)
        fused.attach_gather(g)
        fused.set_inputs(ref[lo:hi], src[lo:hi], tgt_pos=tp[lo:hi])
        fused.step()
        fused.step()   # second step = CUDA-graph replay with the peer pointers baked in
        torch.cuda.synchronize()
        g.barrier()
        assert torch.equal(g.frames, rgb8), f"rank {rank}: fused gather ({g.mode}) differs from the NCCL all-gather"
        assert torch.equal(fused.out["rgb_u8"], rgb8[lo:hi])
        g.barrier()
        if rank == 0:
            print(f"FUSED_OK world={world} mode={g.mode}")
    # one symmetric allocation carved into two gathered buffers (one per frame lane, runtime.FrameGather.slot)
    try:
        g2 = FrameGather(hi - lo, H, W, dev, slots=2)
    except Exception as e:
        g2 = None
        if rank == 0:
            print(f"FUSED_SKIP slots: {type(e).__name__}: {e}")
    if g2 is not None:
        pipes = []
        for k in range(2):
            pk = MSIPipeline(wts, H, W, P, ngf, batch=hi - lo, device=dev)
            pk.attach_gather(g2.slot(k))
            # lane 1 renders the frames in reversed order, so the two slots must differ
            sel = slice(lo, hi) if k == 0 else slice(hi - 1, lo - 1 if lo > 0 else None, -1)
            pk.set_inputs(ref[sel].copy(), src[sel].copy(), tgt_pos=tp[sel].copy())
            pk.step()
            pipes.append(pk)
        torch.cuda.synchronize()
        g2.barrier()
        assert torch.equal(g2.slot(0).frames, rgb8), f"rank {rank}: slot 0 differs from the NCCL all-gather"
        want1 = all_gather_frames(pipes[1].out["rgb_u8"], world)
        assert torch.equal(g2.slot(1).frames, want1), f"rank {rank}: slot 1 differs"
        assert not torch.equal(g2.slot(1).frames, g2.slot(0).frames)
        g2.barrier()
        if rank == 0:
            print(f"FUSED_OK world={world} slots=2 mode={g2.mode}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
