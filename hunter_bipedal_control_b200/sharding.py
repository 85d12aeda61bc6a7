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


SHARD_ID_BYTES = 128


def native_partition(total, world_size, rank):
    """hb_shard_partition: the same block rule behind the C ABI."""
    import ctypes as C
    from . import api
    lo, n = C.c_int(), C.c_int()
    api._check(api.load_library().hb_shard_partition(int(total), int(world_size), int(rank), C.byref(lo), C.byref(n)), "hb_shard_partition")
    return lo.value, lo.value + n.value


def native_sort_by_schedule(mode):
    """hb_shard_sort_by_schedule: (perm, inverse) as int32 arrays; same grouping as sort_by_schedule."""
    from . import api
    mode = np.ascontiguousarray(mode, dtype=np.int32)
    B, nodes = mode.shape
    perm = np.zeros(B, dtype=np.int32); inv = np.zeros(B, dtype=np.int32)
    api._check(api.load_library().hb_shard_sort_by_schedule(B, nodes, api._ptr(mode), api._ptr(perm), api._ptr(inv)), "hb_shard_sort_by_schedule")
    return perm, inv


def unique_id():
    """hb_shard_unique_id (rank 0): the 128 bytes every rank passes to Shard()."""
    import ctypes as C
    from . import api
    buf = (C.c_uint8 * SHARD_ID_BYTES)()
    api._check(api.load_library().hb_shard_unique_id(buf), "hb_shard_unique_id")
    return bytes(buf)


class Shard:
    """Owner of one hb_shard: this rank's block of a job of `total` instances and the NCCL gather of its output rows (C ABI, no torch)."""

    def __init__(self, ctx, uid, world_size, rank, total, max_row_doubles=38):
        import ctypes as C
        from . import api
        self._lib = api.load_library()
        self._ctx = ctx
        self._h = C.c_void_p()
        self.world_size, self.rank, self.total = world_size, rank, total
        idbuf = (C.c_uint8 * SHARD_ID_BYTES).from_buffer_copy(uid) if uid is not None else None
        api._check(self._lib.hb_shard_create(ctx._h, idbuf, int(world_size), int(rank), int(total), int(max_row_doubles), C.byref(self._h)), "hb_shard_create", ctx._h)
        lo, n = C.c_int(), C.c_int()
        api._check(self._lib.hb_shard_block(self._h, C.byref(lo), C.byref(n)), "hb_shard_block")
        self.lo, self.hi = lo.value, lo.value + n.value

    def gather(self, rows, inverse=None):
        """rows: torch cuda tensor [n_local, d] float64 (this block); inverse: int32 cuda tensor or None. Returns the device address of the
        gathered [total, d] rows (instance order, on every rank); complete after wait()."""
        import ctypes as C
        from . import api
        out = C.c_void_p()
        rc = self._lib.hb_shard_gather_dev(self._h, int(rows.shape[1]), api._ptr(rows), api._ptr(inverse), C.byref(out))
        if rc == -6:
            raise api.HunterB200Error("hb_shard_gather_dev: " + self._lib.hb_shard_last_error(self._h).decode())
        api._check(rc, "hb_shard_gather_dev", self._ctx._h)
        return out.value

    def wait(self, block_host=False):
        from . import api
        api._check(self._lib.hb_shard_wait(self._h, 1 if block_host else 0), "hb_shard_wait", self._ctx._h)

    def to_host(self, addr, width):
        """Copy a gathered block (address returned by gather(), after wait()) to a numpy array [total, width]."""
        from cuda.bindings import runtime as cudart
        out = np.empty((self.total, int(width)))
        err, = cudart.cudaMemcpy(out.ctypes.data, int(addr), out.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        if int(err) != 0:
            raise RuntimeError("cudaMemcpy of the gathered rows failed: %s" % (err,))
        return out

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.hb_shard_destroy(self._h)
            self._h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
