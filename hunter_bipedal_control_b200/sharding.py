"""Multi-GPU sharding of independent instances: contiguous block partition, no data-path collective;
the only exchange is the final gather of per-instance outputs to rank 0 (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def partition(total, world_size, rank):
    """Contiguous block [lo, hi) of `total` instances owned by `rank` (block sizes differ by at most one)."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sort_by_schedule(mode):
    """Permutation that groups instances with the same mode pattern (warp-uniform control flow inside a CTA); returns (perm, inverse)."""
    keys = [tuple(np.asarray(m).tolist()) for m in mode]
    perm = np.array(sorted(range(len(keys)), key=lambda i: keys[i]), dtype=np.int64)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    return perm, inv


def gather_to_rank0(local, total, world_size, rank, dist, device=None):
    """Gather per-instance rows (torch tensor [n_local, d]) from every rank to rank 0 in instance order.
    Blocks are padded to the largest block so one all_gather suffices (outputs are tiny: <= 656 B per instance)."""
    import torch
    sizes = [partition(total, world_size, r) for r in range(world_size)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(out, pad)
    if rank != 0:
        return None
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
