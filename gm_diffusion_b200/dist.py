"""Multi-GPU plumbing: the path shards by image — every (prompt, seed) is an independent trajectory, so
ranks take contiguous slices of the global batch, replicate the weights, run with NO collective inside the loop, and
all-gather the outputs once at the end (NCCL over NVLink/NVSwitch; gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `total` images: the first `total % world` ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def gather_outputs(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-rank output slices (possibly ragged by one image) back into global batch order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    max_n = max(b - a for a, b in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], 0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class CfgPair:
    """Optional latency mode of SURVEY.md §8e: two ranks work on the SAME images.  Rank 0 of the pair runs the unconditional half
    of the SDR UNet's classifier-free-guidance batch, rank 1 the conditional half; the two eps tensors (64 KB per image) are
    exchanged with one all-gather per step and both ranks run the fused scheduler step redundantly (deterministic, so their
    latents stay bit-identical).  The GM UNet (no CFG) is split by images across the pair and its eps all-gathered the same way.
    Because every kernel is batch-independent, the result equals the single-GPU result bit for bit."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("CfgPair needs an initialised torch.distributed process group")
        if dist.get_world_size(group) != 2:
            raise ValueError(f"CfgPair works on a group of exactly two ranks, got {dist.get_world_size(group)}")
        self.group = group
        self.rank = dist.get_rank(group)

    def gather(self, local: torch.Tensor, out: torch.Tensor) -> None:
        """out[r] = rank r's `local` (NCCL: stream-ordered on the current CUDA stream).  With the gloo backend — two ranks sharing
        ONE GPU in the single-GPU test leg, where two NCCL ranks cannot coexist and kernels of different processes must never wait
        on one another — the exchange is staged through the host."""
        if dist.get_backend(self.group) == "gloo":
            host = torch.empty(out.shape, dtype=out.dtype)
            dist.all_gather_into_tensor(host, local.cpu(), group=self.group)
            out.copy_(host)
            return
        dist.all_gather_into_tensor(out, local, group=self.group)
