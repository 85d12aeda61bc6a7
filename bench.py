#!/usr/bin/env python3
"""bench.py -- headline benchmark of the batched Hunter NMPC + WBC control step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W              this repo's CUDA path (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K ...    the reference arm: the CPU restatement of the reference's
                                                             algorithm (oracle/, "port": the upstream OCS2+qpOASES binaries
                                                             cannot be built offline) on all host threads, rank 0 only

One "step" = one pass of the hot path over one batch: for every instance one SQP iteration of the N=100, dt=10 ms centroidal
NMPC from the initializer's cold start (LQ approximation, projection, Riccati, forward pass, filter line search), policy
evaluation at t0 + 2 ms, and one WeightedWbc QP. Workload at one GPU = BASELINE.json configs[1]: 1024 instances, trot gait,
randomised initial base pose (seed 20240901 + instance index); N>1 keeps 1024 instances per GPU (weak scaling), sharded by
contiguous blocks with no data-path collective; the per-instance torques are gathered to rank 0 with NCCL every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MPC+WBC control-step solves/sec (Hunter N=100)"
UNIT = "solves/s"
PER_GPU_BATCH = 1024
HORIZON_N, DT = 100, 0.01
T_POLICY = 0.002
# SURVEY.md 8(d) / BASELINE.md 4: algorithmic HBM bytes and flops of one control-step solve
BYTES_PER_MPC_SOLVE = 0.48e6
BYTES_PER_WBC_SOLVE = 616 + 304
FLOPS_PER_SOLVE = 3.6e7 + 2.0e6
FP64_NOMINAL_TFLOPS = 37.0   # B200 FP64 CUDA-core peak (not in MEASURED_PEAKS.json; nominal)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index; self.samples = []; self.reasons = set(); self.max_mhz = None; self._stop = False; self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.15)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        self._stop = True
        if self._t:
            self._t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def workload(B, seed_offset=0, with_compact=False):
    from hunter_bipedal_control_b200 import scenarios as S
    x0, x_ref, swing, mode, compacts = S.make_batch(B, HORIZON_N, DT, gait="trot", cmd_vel=(0.2, 0.0, 0.0, 0.0), seed=20240901 + seed_offset, return_compact=True)
    rbd = S.consistent_rbd(x0, np.random.default_rng(20240901 + seed_offset), 0.0)
    if with_compact:
        return x0, x_ref, swing, mode, rbd, S.pack_references(compacts, 2 * HORIZON_N * DT)
    return x0, x_ref, swing, mode, rbd


def cpu_control_steps(x0, x_ref, swing, mode, rbd, threads):
    """The oracle's control step on host threads; returns (seconds, torques)."""
    from oracle import hbo
    n = x0.shape[0]
    xt = np.zeros((n, HORIZON_N + 1, 22)); ut = np.zeros((n, HORIZON_N, 22))
    for i in range(n):
        xt[i], ut[i] = hbo.mpc_cold_start(HORIZON_N, DT, x0[i], mode[i])
    t = time.perf_counter()
    xt1, ut1, _ = hbo.mpc_iteration_batch(HORIZON_N, DT, x0, x_ref, swing, mode, xt, ut, threads=threads)
    al = T_POLICY / DT
    xd = (1 - al) * xt1[:, 0] + al * xt1[:, 1]; ud = (1 - al) * ut1[:, 0] + al * ut1[:, 1]
    sol, _ = hbo.wbc_solve_batch(xd, ud, rbd, mode[:, 0], np.zeros(n, dtype=np.uint8), 1e-8, threads=threads)
    return time.perf_counter() - t, sol[:, 28:]


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = max(2 * cores, 8)
    data = workload(n)
    for _ in range(args.warmup):
        cpu_control_steps(*[d[:min(n, cores)] for d in data], threads=cores)
    ts = []
    for _ in range(args.steps):
        dt_, _ = cpu_control_steps(*data, threads=cores)
        ts.append(dt_)
    total = float(np.sum(ts))
    val = n * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "configs[1]: Hunter trot, N=100 dt=10 ms, randomised initial base pose; %d-instance sample per step" % n,
                       "note": "CPU restatement of the reference algorithm (oracle port); upstream OCS2+qpOASES binaries cannot be built offline"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d control steps per bench step, all host threads" % n},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    import torch
    import hunter_bipedal_control_b200 as hb
    from hunter_bipedal_control_b200 import sharding
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = args.batch
    total_B = B * world
    lo, hi = sharding.partition(total_B, world, rank)
    x0, x_ref, swing, mode, rbd, refs = workload(B, seed_offset=lo, with_compact=True)     # instance index = global index
    ctx = hb.Context(horizon_N=HORIZON_N, dt=DT, max_batch=B, device=local)
    stream = torch.cuda.ExternalStream(ctx.stream_handle, device=dev)
    to = lambda a, dt_=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt_)
    d_x0, d_xref, d_swing, d_rbd = to(x0), to(x_ref), to(swing), to(rbd)
    d_mode = to(mode, torch.int32)
    d_xt0 = torch.zeros((B, HORIZON_N + 1, 22), dtype=torch.float64, device=dev); d_ut0 = torch.zeros((B, HORIZON_N, 22), dtype=torch.float64, device=dev)
    ctx.mpc_cold_start_dev(d_x0, d_mode, d_xt0, d_ut0)
    ctx.sync()
    d_xt = d_xt0.clone(); d_ut = d_ut0.clone()
    d_info = torch.zeros((B, 7), dtype=torch.float64, device=dev)
    d_sol = torch.zeros((B, 38), dtype=torch.float64, device=dev); d_tau = torch.zeros((B, 10), dtype=torch.float64, device=dev)
    d_st = torch.zeros(B, dtype=torch.int32, device=dev)
    gathered = None

    def step_device():
        nonlocal gathered
        with torch.cuda.stream(stream):
            d_xt.copy_(d_xt0, non_blocking=True); d_ut.copy_(d_ut0, non_blocking=True)   # every step starts from the initializer's cold start
        ctx.control_step_dev(T_POLICY, d_x0, d_xref, d_swing, d_mode, d_rbd, d_xt, d_ut, d_info, d_sol, d_tau, d_st)
        if world > 1:
            with torch.cuda.stream(stream):
                gathered = sharding.gather_to_rank0(d_tau, total_B, world, rank, dist)

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync(); torch.cuda.synchronize(dev)

    # ---------------- device-resident measurement
    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.profile_enable(True)
    l0 = ctx.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - l0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    # ---------------- end-to-end through the host-pointer C ABI, pinned host buffers, copies inside the timed region.
    # e2e      : hb_resident_cycle_batch -- the closed-loop call: t0 / x0 / compact references / rbd in, info / WBC solution / torques out;
    #            reference expansion and the initializer cold start run on the device, the primal solution stays resident.
    # e2e_full : hb_control_step_batch -- node-sampled references and the full state / input trajectories cross PCIe both ways.
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_x0, h_xref, h_swing, h_rbd, h_mode = pin(x0), pin(x_ref), pin(swing), pin(rbd), pin(mode)
    import ctypes as C
    ref_bytes = C.sizeof(refs)
    h_refs = torch.empty(ref_bytes, dtype=torch.uint8).pin_memory()
    C.memmove(h_refs.data_ptr(), C.addressof(refs), ref_bytes)
    h_t0 = torch.zeros(B, dtype=torch.float64).pin_memory()
    h_info = torch.zeros((B, 7), dtype=torch.float64).pin_memory(); h_sol = torch.zeros((B, 38), dtype=torch.float64).pin_memory()
    h_tau = torch.zeros((B, 10), dtype=torch.float64).pin_memory(); h_st = torch.zeros(B, dtype=torch.int32).pin_memory()
    lib = hb.load_library()
    P = lambda t: C.c_void_p(t.data_ptr())

    def step_e2e():
        rc = lib.hb_resident_cycle_batch(ctx._h, B, 1, C.c_double(T_POLICY), P(h_t0), P(h_x0), P(h_refs), P(h_rbd), P(h_info), P(h_sol), P(h_tau), P(h_st))
        assert rc == 0, rc

    for _ in range(2):
        step_e2e()
    barrier()
    _td = d_tau.cpu().numpy()
    e2e_tau_diff = float(np.abs(h_tau.numpy() - _td).max() / max(1e-300, np.abs(_td).max()))     # same work as the device-resident step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = (h_t0.numel() + h_x0.numel() + h_rbd.numel()) * 8 + ref_bytes
    d2h = (h_info.numel() + h_sol.numel() + h_tau.numel()) * 8 + h_st.numel() * 4
    # e2e_plan_cycle: hb_resident_plan_cycle_batch -- plan inputs (t0, x0, cmd_vel, gait; 352 B per instance) and rbd in; foot positions, the
    # reference planner (gait tiling, swing planner, IK joint references), expansion, cold start, solve, WBC all on the device.
    ins = hb.make_plan_inputs(np.zeros(B), HORIZON_N * DT, x0, (0.2, 0.0, 0.0, 0.0), None, "trot", 0.1)
    plan_bytes = C.sizeof(ins)
    h_ins = torch.empty(plan_bytes, dtype=torch.uint8).pin_memory()
    C.memmove(h_ins.data_ptr(), C.addressof(ins), plan_bytes)
    h_ps = torch.zeros(B, dtype=torch.int32).pin_memory()

    def step_plan():
        rc = lib.hb_resident_plan_cycle_batch(ctx._h, B, 1, C.c_double(T_POLICY), P(h_ins), P(h_rbd), P(h_info), P(h_sol), P(h_tau), P(h_st), P(h_ps))
        assert rc == 0, rc

    for _ in range(2):
        step_plan()
    barrier()
    plan_ok = bool((h_ps.numpy() == 0).all() and (h_st.numpy() == 0).all())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_plan()
    barrier()
    plan_s = time.perf_counter() - t0
    plan_h2d = plan_bytes + h_rbd.numel() * 8
    plan_d2h = (h_info.numel() + h_sol.numel() + h_tau.numel()) * 8 + h_st.numel() * 4 + h_ps.numel() * 4
    # every full-trajectory step gets its own pre-initialised in/out trajectory buffers (cold start): no host-side reset in the timed region
    n_e2e = args.steps + 2
    h_xts = [d_xt0.cpu().pin_memory() for _ in range(n_e2e)]; h_uts = [d_ut0.cpu().pin_memory() for _ in range(n_e2e)]
    h_xt, h_ut = h_xts[0], h_uts[0]
    e2e_i = [0]

    def step_e2e_full():
        xt_, ut_ = h_xts[e2e_i[0]], h_uts[e2e_i[0]]
        e2e_i[0] += 1
        rc = lib.hb_control_step_batch(ctx._h, B, C.c_double(T_POLICY), P(h_x0), P(h_xref), P(h_swing), P(h_mode), P(h_rbd), P(xt_), P(ut_), P(h_info),
                                       P(h_sol), P(h_tau), P(h_st))
        assert rc == 0, rc

    for _ in range(2):
        step_e2e_full()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_full()
    barrier()
    full_s = time.perf_counter() - t0
    full_h2d = (h_x0.numel() + h_xref.numel() + h_swing.numel() + h_rbd.numel() + h_xt.numel() + h_ut.numel()) * 8 + h_mode.numel() * 4
    full_d2h = (h_xt.numel() + h_ut.numel() + h_info.numel() + h_sol.numel() + h_tau.numel()) * 8 + h_st.numel() * 4
    clocks = sampler.stop() if rank == 0 else None
    # ---------------- max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3, full_s * 1e3, plan_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, full_ms, plan_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    else:
        e2e_ms, full_ms, plan_ms = e2e_s * 1e3, full_s * 1e3, plan_s * 1e3
    ok = bool((d_st == 0).all().item())
    if rank == 0:
        peaks, src = load_peaks()
        value = total_B * args.steps / (ms * 1e-3)
        e2e_val = total_B * args.steps / (e2e_ms * 1e-3)
        mpc_names = [n for n in prof if n.startswith("mpc_")]
        top = max(mpc_names, key=lambda n: prof[n]["ms"])     # dominant kernel of the step
        bk = prof[top]
        bk_ms = bk["ms"] / max(1, bk["launches"])
        ach = B * BYTES_PER_MPC_SOLVE / (bk_ms * 1e-3) / 1e9 if bk_ms > 0 else None
        step_ms = ms / args.steps
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (same workload, 1024 instances)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            kname = {"mpc_lq_project": "lq_kernel", "mpc_linearise": "lin_kernel", "mpc_riccati": "riccati_kernel", "mpc_forward_linesearch": "forward_linesearch2_kernel"}[top]
            traffic = tj[kname]["dram_bytes"] * B / tj[kname]["instances"]
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "configs[1]: %d Hunter instances per GPU, trot gait, N=100 dt=10 ms, randomised initial base pose (seed 20240901+i), "
                           "one SQP iteration from the initializer cold start + policy eval at 2 ms + WeightedWbc QP" % B,
                           "instances_total": total_B, "parallelism": "instances sharded in contiguous blocks, NCCL gather of torques" if world > 1 else "single GPU",
                           "l2": "per-step working set (node records %.0f MB + references/trajectories %.0f MB) exceeds the 126 MB L2" % (B * HORIZON_N * (1200 + 2320 + 368) * 8 / 1e6, full_h2d / 1e6)},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps,
                        "call": "hb_resident_cycle_batch(cold_start=1): t0, x0, compact references, rbd in; info, WBC solution, torques, status out",
                        "torque_max_rel_diff_vs_device_path": e2e_tau_diff},
                "e2e_plan_cycle": {"value": total_B * args.steps / (plan_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(plan_h2d), "d2h_bytes_per_step": int(plan_d2h),
                                   "ms_per_step": plan_ms / args.steps, "all_planned_and_solved": plan_ok,
                                   "call": "hb_resident_plan_cycle_batch(cold_start=1): plan inputs + rbd in; reference planner (P1, P3, P4, P5) on the device; same gait / command as the workload"},
                "e2e_full_trajectories": {"value": total_B * args.steps / (full_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(full_h2d),
                                          "d2h_bytes_per_step": int(full_d2h), "ms_per_step": full_ms / args.steps,
                                          "call": "hb_control_step_batch: node-sampled references and full trajectories both ways"},
                "gpu_launches": int(launches),
                "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
                "roofline": {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": (ach / peaks["hbm_gbs"]) if ach else None, "peak_source": src, "traffic": traffic,
                             "share_of_step": bk_ms / step_ms if step_ms > 0 else None,
                             "note": "latency/FP64-bound path: the HBM fraction is small by construction (SURVEY 8d); see roofline_fp64"},
                "roofline_fp64": {"achieved_tflops": value / world * FLOPS_PER_SOLVE / 1e12, "peak_tflops_nominal": FP64_NOMINAL_TFLOPS,
                                  "frac": value / world * FLOPS_PER_SOLVE / 1e12 / FP64_NOMINAL_TFLOPS},
                "clocks": clocks, "all_converged": ok}
        if not args.no_cpu_baseline and world == 1:     # reported on rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            n = max(cores, 4)
            data = [d[:n] for d in (x0, x_ref, swing, mode, rbd)]
            t_all, tau_cpu = cpu_control_steps(*data, threads=cores)
            t_one, _ = cpu_control_steps(*[d[:2] for d in data], threads=1)
            err = float(np.abs(tau_cpu - d_tau[:n].cpu().numpy()).max() / max(1.0, np.abs(tau_cpu).max()))
            line["cpu_baseline"] = {"value": n / t_all, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d control steps of the same workload, all host threads; single-thread %.2f solves/s" % (n, 2 / t_one),
                                    "single_thread_value": 2 / t_one, "torque_rel_err_vs_gpu": err}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
