"""Worker of tests/test_gpu_shard.py (one process per GPU, launched with torch.distributed.run): the WBC-only control work of a job of
`total` mixed-schedule instances, sharded through the C-ABI shard API (hb_shard_*): native partition, schedule sort inside the shard, solve,
un-permute + NCCL all-gather behind the compute stream. Rank 0 solves the WHOLE job on its own GPU as well and compares bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import hunter_bipedal_control_b200 as hb                            # noqa: E402
from hunter_bipedal_control_b200 import scenarios as sc, sharding   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    total = int(sys.argv[1])
    dist.init_process_group("gloo")                                  # only carries the 128-byte communicator id
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    uid = [sharding.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = hb.Context(horizon_N=1, dt=0.01, max_batch=total, device=local)
    shard = sharding.Shard(ctx, uid[0], world, rank, total, max_row_doubles=38)
    assert (shard.lo, shard.hi) == sharding.partition(total, world, rank)
    # the whole job, generated identically on every rank
    rng = np.random.default_rng(11)
    x = sc.random_initial_states(total, seed=77)
    u = np.zeros((total, 22)); u[:, 2] = u[:, 5] = u[:, 8] = u[:, 11] = sc.TOTAL_MASS * 9.81 / 4
    rbd = sc.consistent_rbd(x, rng, 0.05)
    mode = rng.choice([3, 3, 1, 2], size=total).astype(np.int32)
    lo, hi = shard.lo, shard.hi
    perm, inv = sharding.native_sort_by_schedule(mode[lo:hi, None])
    to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_x, d_u, d_rbd, d_m = to_dev(x[lo:hi][perm]), to_dev(u[lo:hi][perm]), to_dev(rbd[lo:hi][perm]), to_dev(mode[lo:hi][perm])
    d_inv = to_dev(inv)
    n = hi - lo
    d_sol = torch.zeros((n, 38), dtype=torch.float64, device=dev); d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    lib, P = ctx._lib, hb.api._ptr
    results = []
    for rep in range(3):                                              # three back-to-back gathers: the double buffering is exercised
        hb.api._check(lib.hb_wbc_solve_batch_dev(ctx._h, n, P(d_x), P(d_u), P(d_rbd), P(d_m), None, P(d_sol), P(d_st)), "hb_wbc_solve_batch_dev", ctx._h)
        addr_tau = shard.gather(d_sol[:, 28:].contiguous(), d_inv)   # torque rows (10 doubles), sorted order -> instance order
        shard.wait(block_host=True)
        results.append(shard.to_host(addr_tau, 10))
    addr_sol = shard.gather(d_sol, d_inv)                            # a wider row through the same shard
    shard.wait(block_host=True)
    sol_all = shard.to_host(addr_sol, 38)
    assert np.array_equal(results[0], results[1]) and np.array_equal(results[1], results[2])
    assert np.array_equal(sol_all[:, 28:], results[0])
    # every rank holds the same gathered block (all-gather)
    digest = torch.tensor([float(np.abs(results[0]).sum())], dtype=torch.float64)
    lst = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(lst, digest)
    assert all(float(t) == float(digest) for t in lst)
    if rank == 0:
        sol_ref, st_ref = ctx.wbc_solve(x, u, rbd, mode)             # the whole job on one GPU, instance order
        assert (st_ref == 0).all()
        assert np.array_equal(sol_ref, sol_all), np.abs(sol_ref - sol_all).max()
        print("SHARD_OK world=%d total=%d blocks=%s" % (world, total, [sharding.partition(total, world, r) for r in range(world)]))
    shard.close(); ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
