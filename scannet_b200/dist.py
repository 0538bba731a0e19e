"""Multi-GPU plumbing for the scene-sharded path: one process per GPU, one scene per rank,
NO data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only for the barrier and
the end-of-run reduction of (frames, device-measured seconds)."""
from __future__ import annotations

import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def scene_seed_for_rank(rank: int, base_seed: int = 0) -> int:
    """scene i -> rank i mod world: with one scene per rank the scene id is the rank"""
    return base_seed + rank


class Group:
    def __init__(self, backend: str | None = None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = env_rank()
        self.device = device
        self.backend = backend
        if self.world > 1:
            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            dist.init_process_group(backend or "gloo", **kw)

    def barrier(self):
        if self.device is not None and self.device.type == "cuda":
            self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        if self.device is not None and self.device.type == "cuda":
            self.torch.cuda.synchronize()

    def reduce_throughput(self, frames: int, ms: float):
        """returns (total frames over all ranks, max ms over ranks) — value = total / max"""
        t = self.torch.tensor([float(frames), 0.0], dtype=self.torch.float64, device=self.device or "cpu")
        m = self.torch.tensor([float(ms)], dtype=self.torch.float64, device=self.device or "cpu")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX)
        return int(t[0].item()), float(m.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
