"""Column sharding of the unitary across ranks -- host logic only.

Every gate acts on row bits, so the 2^n columns of U are independent: rank r of W owns the
columns [r * 2^n / W, (r+1) * 2^n / W), runs the SAME gate stream on its shard with no
communication, and the equivalence check needs ONE all-reduce (SUM) of 8 doubles.
One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing."""
from __future__ import annotations

import numpy as np


def shard_columns(n_qubits: int, rank: int, world: int) -> tuple[int, int]:
    """(col0, ncols) of this rank's shard; world must be a power of two <= 2^n."""
    dim = 1 << n_qubits
    if world < 1 or world & (world - 1) or world > dim:
        raise ValueError(f"world size {world} must be a power of two <= 2^{n_qubits}")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    ncols = dim // world
    return rank * ncols, ncols


def allreduce_sums(sums, device=None) -> np.ndarray:
    """Adds the 8 equivalence sums over all ranks with one all_reduce; a no-op when
    torch.distributed is not initialised (single GPU)."""
    sums = np.asarray(sums, dtype=np.float64)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return sums
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sums
    t = torch.from_numpy(sums.copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def max_over_ranks(value: float, device=None) -> float:
    try:
        import torch
        import torch.distributed as dist
    except ImportError:  # pragma: no cover
        return value
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
