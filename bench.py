#!/usr/bin/env python3
"""bench.py -- MPC solves/sec of the batched SQP-RTI hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA)
  python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm (oracle port, all host cores)
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU, env-sharded

A "step" is ONE SQP-RTI solve (robot_ocp_problem.py:195) of a batch of independent synthetic
scenarios (4096 random start/obstacle scenarios, N=20, Tf=2 s, 5 moving
obstacles) per GPU.  Every step starts from the same cold iterate (set_initial_guess state), so
every step does identical work; the iterate is restored from a pristine device copy and L2 is
flushed between steps, both outside the per-step CUDA-event pairs.

Prints one JSON line (rank 0).  See README / DESIGN.md for the field definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "dynamic-obstacle-avoidance-mpc_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "MPC solves/sec (SQP-RTI, N=20, Tf=2 s, 5 obstacles)"
UNIT = "solves/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=4096)
    ap.add_argument("--N", type=int, default=20)
    ap.add_argument("--n-obst", type=int, default=5)
    ap.add_argument("--tf", type=float, default=2.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 2 warp-per-problem (node-private rows in global memory), 3 warp-per-problem all-shared")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU sample (0 = auto)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip BASELINE configs 1/3/4/5 (extra keys of the JSON line)")
    return ap.parse_args()


def workload_name(a):
    return (f"batched SQP-RTI, {a.batch_per_gpu} random start/obstacle scenarios per GPU, N={a.N}, "
            f"Tf={a.tf:g} s, {a.n_obst} moving obstacles, cold start (set_initial_guess iterate)")


def config_dict(a, n_gpus):
    """the SAME dictionary in both arms (b200 / reference) for the same command line: what is measured, not how"""
    return {"workload": workload_name(a), "global_batch": int(a.batch_per_gpu * n_gpus),
            "l2": "GPU arm: flushed between steps (256 MiB memset), both outside the per-step event pairs; CPU arm: not applicable",
            "kernel": {0: "auto", 2: "warp-per-problem", 3: "warp-per-problem (all shared)"}.get(a.kernel, str(a.kernel)),
            "parallelism": f"env-sharded x{n_gpus}, no data-path collective"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_arm(a, sample, nthreads, reps=1):
    """times the oracle port on `sample` problems of the same workload with `nthreads` host threads"""
    from b200mpc import world
    from oracle import pyoracle as po
    dims = world.MpcDims(a.N, a.n_obst, a.tf)
    sc = world.random_scenarios(sample, dims, seed=a.seed)
    cfg = po.default_config(a.N, a.n_obst, a.tf)
    best = None
    for _ in range(reps):
        x, u, sl, su = (sc[k].copy() for k in ("x", "u", "sl", "su"))
        t0 = time.perf_counter()
        st, it, _ = po.rti_solve_batch(cfg, sc["x0"], sc["p"], sc["yref"], sc["zl"], sc["Zl"], x, u, sl, su, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, best, float(it.mean())


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port; acados itself is not installable
    offline -- parity unpinned) on all host cores; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = a.cpu_sample or min(a.batch_per_gpu, 4096)
    for _ in range(a.warmup):
        cpu_arm(a, min(sample, 256), cores)
    times = []
    for _ in range(a.steps):
        v, dt, mean_it = cpu_arm(a, sample, cores)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(a, a.gpus),
        "sample_per_step": sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} problems of the same workload per step, pthreads over problems, "
                                   "oracle/mpc_oracle.c (CPU restatement of the acados algorithm -- not acados)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_qp_iter": mean_it,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist
    from b200mpc import BatchedMpcSolver, dfma_peak_tflops, flops, sharding, world

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = a.batch_per_gpu
    dims = world.MpcDims(a.N, a.n_obst, a.tf)
    # env-sharded: rank r owns the contiguous chunk [r*B, (r+1)*B) of ONE global seeded scenario set;
    # no cross-GPU traffic until the final reduction of a few scalars
    sc = sharding.shard_dict(world.random_scenarios(B * world_size, dims, seed=a.seed), rank, world_size)
    solver = BatchedMpcSolver(a.N, a.n_obst, a.tf, max_batch=B, device=local_rank, kernel=a.kernel)
    names_in = ("x0", "p", "yref", "zl", "Zl")
    names_it = ("x", "u", "sl", "su")
    dev_t = {k: torch.from_numpy(sc[k]).to(dev) for k in names_in + names_it}
    for k in names_in:
        solver.set(k, dev_t[k])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def restore():
        for k in names_it:
            solver.set(k, dev_t[k])

    stream = torch.cuda.current_stream(dev)

    def step_device():
        restore(); flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.solve(B)
        e1.record(stream)
        return e0, e1

    for _ in range(max(a.warmup, 3)):
        step_device()
    torch.cuda.synchronize(dev)
    peak_tf = dfma_peak_tflops(local_rank)

    # ---------------- timed region: device-resident inputs ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    l0 = solver.launch_count
    evs = [step_device() for _ in range(a.steps)]
    torch.cuda.synchronize(dev)
    if world_size > 1:
        dist.barrier()
    l1 = solver.launch_count
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = float(np.sum(step_ms))
    qp_iter = solver.get_int("qp_iter", B).cpu().numpy()
    status = solver.get_int("status", B).cpu().numpy()
    u_device_path = solver.get("u", B).cpu().numpy()          # result of the device-resident path (compared with the host path below)
    # restore()'s 4 layout kernels per step are not the solve; count the solve launches only
    gpu_launches = (l1 - l0) - 4 * a.steps

    # ---------------- e2e: host buffers through the public API ----------------
    host = {k: torch.from_numpy(sc[k]).pin_memory().numpy() for k in names_in}
    host_it0 = {k: sc[k].copy() for k in names_it}
    host_it = {k: torch.from_numpy(sc[k].copy()).pin_memory().numpy() for k in names_it}

    def step_host():
        for k in names_it:
            np.copyto(host_it[k], host_it0[k])
        t0 = time.perf_counter()
        st, it = solver.solve_host(host["x0"], host["p"], host["yref"], host["zl"], host["Zl"],
                                   host_it["x"], host_it["u"], host_it["sl"], host_it["su"])
        return time.perf_counter() - t0

    for _ in range(2):
        step_host()
    if world_size > 1:
        dist.barrier()
    e2e_steps = max(3, min(a.steps, 10))
    e2e_s = float(np.sum([step_host() for _ in range(e2e_steps)]))
    h2d = sum(host[k].nbytes for k in names_in) + sum(host_it[k].nbytes for k in names_it)
    d2h = sum(host_it[k].nbytes for k in names_it) + 2 * 4 * B
    # the host path (fresh H2D copies of every input) reproduces the device-resident path bit for bit
    assert np.array_equal(host_it["u"], u_device_path)

    # ---------------- steady state: the same solve on WARM iterates ----------------
    # 50 closed-loop control steps on device (obstacle prediction, slack weights, solve, plant/obstacle step, shift:
    # robot_ocp_problem.py:184-260), then the solve of control step 51 is timed like the cold one
    K_WARM = 50
    solver.set("x0", dev_t["x0"]); solver.set("obst", torch.from_numpy(sc["obst"]).to(dev)); solver.set("goal", torch.from_numpy(sc["goal"]).to(dev))
    solver.set("yref", dev_t["yref"])
    restore(); solver.loop_reset(B)
    solver.closed_loop(B, K_WARM, random_move=False, init_guess_when_error=True, tol=-1.0)   # tol < 0: nobody "reaches the goal"
    solver.obst_predict(B); solver.prepare_slack(B)
    warm = {k: solver.get(k, B).clone() for k in names_it}

    def step_warm():
        for k in names_it:
            solver.set(k, warm[k])
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream); solver.solve(B); e1.record(stream)
        return e0, e1

    for _ in range(2):
        step_warm()
    torch.cuda.synchronize(dev)
    if world_size > 1:
        dist.barrier()
    wevs = [step_warm() for _ in range(a.steps)]
    torch.cuda.synchronize(dev)
    warm_ms = float(np.sum([e0.elapsed_time(e1) for e0, e1 in wevs]))
    warm_it = float(solver.get_int("qp_iter", B).sum().item())
    kernel_name = solver.kernel_name

    # ---------------- on-device closed loop (b2mpc_closed_loop): env-steps/s against the warm solve rate ----------------
    def closed_loop_rate(s, nb, steps, sc_):
        s.set("x0", torch.from_numpy(sc_["x0"]).to(dev)); s.set("obst", torch.from_numpy(sc_["obst"]).to(dev))
        s.set("goal", torch.from_numpy(sc_["goal"]).to(dev)); s.set("yref", torch.from_numpy(sc_["yref"]).to(dev))
        s.loop_reset(nb); s.reset(nb)
        s.closed_loop(nb, 10, random_move=True, seed=1, tol=-1.0)           # warm-up: leaves warm-started problems behind
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        # chunks of 8 control steps per call; the iteration count of the last step of every chunk is sampled on the device
        # (no synchronisation inside the timed region) so that the rate can be compared per interior-point iteration
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        e0.record(stream)
        for _ in range(steps // 8):
            s.closed_loop(nb, 8, random_move=True, seed=1, tol=-1.0)
            acc += s.get_int("qp_iter", nb).sum()
        e1.record(stream); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / (8 * (steps // 8)), float(acc.item()) / (nb * (steps // 8))

    cl_ms, cl_it = closed_loop_rate(solver, B, 40, sc)

    # ---------------- p50 latency of ONE problem (the reference's one-at-a-time use) ----------------
    lat1, lat_kernel_name, lat_detail = None, None, None
    if rank == 0:
        # one problem at a time is the latency kernel's job (CTA-per-problem, mpc_cta.cuh) where it is instantiated
        from b200mpc import latency_kernel
        s1 = BatchedMpcSolver(a.N, a.n_obst, a.tf, max_batch=1, device=local_rank, kernel=a.kernel or latency_kernel(a.N, a.n_obst))
        lat_kernel_name = s1.kernel_name
        # p50 over SOLVES: the first 48 scenarios of the workload, each solved alone (B = 1), three times round; all solves are
        # enqueued back to back with one event pair around each and ONE synchronisation at the end (the GPU never idles, so
        # its clocks stay where the throughput numbers saw them).  The first round is warm-up.
        n_lat = min(48, B)
        pairs, its1 = [], []
        for rep in range(3):
            for i in range(n_lat):
                for k in names_in + names_it:
                    s1.set(k, dev_t[k][i:i + 1])
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); s1.solve(1); e1.record(stream)
                if rep > 0:
                    pairs.append((e0, e1))
                if rep == 2:
                    its1.append(s1.get_int("qp_iter", 1))
        torch.cuda.synchronize(dev)
        lat_all = np.array([e0.elapsed_time(e1) for e0, e1 in pairs])
        lat1 = float(np.median(lat_all))
        lat_detail = {"p50_ms": lat1, "p90_ms": float(np.quantile(lat_all, 0.9)), "min_ms": float(lat_all.min()), "max_ms": float(lat_all.max()),
                      "problems": n_lat, "median_qp_iter": float(np.median([int(t.item()) for t in its1])),
                      "scenario_0_ms": float(np.median(lat_all[0::n_lat])),
                      "what": "device time of b2mpc_solve(batch = 1) (one launch: a single problem skips the queue kernels), median over the first "
                              f"{n_lat} scenarios of the workload x 2 timed rounds"}
        # BASELINE config 1: one closed-loop episode (B = 1), ms per control step through the fused entry
        sc1 = {k: np.ascontiguousarray(v[:1]) for k, v in sc.items()}
        cfg1_ms, cfg1_it = closed_loop_rate(s1, 1, 100, sc1)
        s1.close()

    # ---------------- BASELINE configs 3 / 4 / 5 (extra keys; the headline above stays configs[1]) ----------------
    extra = {}
    if not a.no_extra_configs:
        solver.close()
        del dev_t, warm

        def timed_cold_solves(N_, M_, tf_, nb, reps, seed):
            d_ = world.MpcDims(N_, M_, tf_)
            sc_ = sharding.shard_dict(world.random_scenarios(nb * world_size, d_, seed=seed), rank, world_size)
            s_ = BatchedMpcSolver(N_, M_, tf_, max_batch=nb, device=local_rank)
            t_ = {k: torch.from_numpy(sc_[k]).to(dev) for k in names_in + names_it}
            for k in names_in:
                s_.set(k, t_[k])
            ms_ = []
            for i in range(reps + 1):
                for k in names_it:
                    s_.set(k, t_[k])
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream); s_.solve(nb); e1.record(stream); torch.cuda.synchronize(dev)
                if i > 0:
                    ms_.append(e0.elapsed_time(e1))
            it_ = float(s_.get_int("qp_iter", nb).float().sum().item())
            name_ = s_.kernel_name
            s_.close()
            return float(np.mean(ms_)), it_, name_

        def cfg_entry(N_, M_, nb, ms_, it_sum, what):
            sm_, mx_ = sharding.reduce_stats({"B": nb, "it": it_sum}, {"ms": ms_}, device=dev)
            mit = sm_["it"] / sm_["B"]
            fl_ = flops.flops_per_solve(N_, M_, mit) * nb
            tf_ach = fl_ / (ms_ * 1e-3) / 1e12
            return {"what": what, "value": sm_["B"] / (mx_["ms"] * 1e-3), "unit": UNIT, "ms_per_step": mx_["ms"], "global_batch": int(sm_["B"]),
                    "mean_qp_iter": mit,
                    "roofline": {"bound": "fp64", "achieved": tf_ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf_ach / peak_tf if peak_tf else None}}

        nb3 = 262144 // world_size
        ms3, it3, _ = timed_cold_solves(a.N, a.n_obst, a.tf, nb3, 2, a.seed)
        extra["config3"] = cfg_entry(a.N, a.n_obst, nb3, ms3, it3, f"262144 scenarios env-sharded over {world_size} GPU(s): {nb3} per GPU, N={a.N}, {a.n_obst} obstacles, cold start")
        ms4, it4, _ = timed_cold_solves(100, 20, 5.0, 4096, 2, a.seed)
        extra["config4"] = cfg_entry(100, 20, 4096, ms4, it4, "long horizon N=100, Tf=5 s, 20 dynamic obstacles, 4096 scenarios per GPU, cold start")
        # config 5: RL loop (src/train.py API) over 65536 parallel environments in total, 5 MPC control steps per env step
        from b200mpc.rl_env import BatchedMpcEnv
        nb5 = 65536 // world_size
        env = BatchedMpcEnv(nb5, mpc_steps=5, max_env_steps=1000, seed=3 + rank, device=local_rank)
        env.reset()
        env.step(env.goal)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n_env_steps = 3
        for _ in range(n_env_steps):
            env.step(env.goal)
        torch.cuda.synchronize(dev)
        dt5 = (time.perf_counter() - t0) / n_env_steps
        it5 = float(env.ocp.solver.get_int("qp_iter", nb5).float().sum().item())
        env.close()
        sm5, mx5 = sharding.reduce_stats({"B": nb5, "it": it5}, {"s": dt5}, device=dev)
        mit5 = sm5["it"] / sm5["B"]
        fl5 = flops.flops_per_solve(20, 5, mit5) * nb5 * 5
        extra["config5"] = {"what": f"RL loop (train.py API: reset/step, action = sub-goal) over {int(sm5['B'])} parallel environments, "
                                    "5 MPC control steps per env step, wall clock incl. observation D2H and Python",
                            "value": sm5["B"] * 5 / mx5["s"], "unit": UNIT, "env_steps_per_s": sm5["B"] / mx5["s"], "global_batch": int(sm5["B"]),
                            "mean_qp_iter": mit5,
                            "roofline": {"bound": "fp64", "achieved": fl5 / dt5 / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                                         "frac": fl5 / dt5 / 1e12 / peak_tf if peak_tf else None}}

    # ---------------- reduce over ranks (max time; sum of work) ----------------
    sums, maxs = sharding.reduce_stats({"B": B, "qp_iter": float(qp_iter.sum()), "fail": float((status != 0).sum()), "warm_it": warm_it},
                                       {"total_ms": total_ms, "e2e_s": e2e_s, "warm_ms": warm_ms, "cl_ms": cl_ms}, device=dev)
    total_ms_max, e2e_s_max = maxs["total_ms"], maxs["e2e_s"]
    tot_B, tot_it, tot_fail = sums["B"], sums["qp_iter"], sums["fail"]

    if rank == 0:
        ms_per_step = total_ms_max / a.steps
        value = tot_B * a.steps / (total_ms_max * 1e-3)
        e2e_value = tot_B * e2e_steps / e2e_s_max
        mean_it = tot_it / tot_B
        fl = flops.flops_per_solve(a.N, a.n_obst, mean_it) * B            # per launch (per GPU)
        kern_s = (float(np.mean(step_ms)) * 1e-3)
        achieved_tf = fl / kern_s / 1e12
        cnt = flops.counts(a.N, a.n_obst)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None                       # dram__bytes_read+write of the solve kernel, one ncu --set full capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr.get("kernel") == kernel_name and tr.get("batch") == B and (a.N, a.n_obst) == (tr.get("N"), tr.get("M")):
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(a, world_size),
            "kernel_resolved": kernel_name,
            "p50_latency_ms_per_batch": float(np.median(step_ms)), "p50_latency_ms_single_problem": lat1,
            "single_problem_kernel": lat_kernel_name, "single_problem_latency": lat_detail,
            "mean_qp_iter": mean_it, "qp_failures": int(tot_fail),
            "steady_state": {"value": tot_B * a.steps / (maxs["warm_ms"] * 1e-3), "unit": UNIT, "ms_per_step": maxs["warm_ms"] / a.steps,
                             "mean_qp_iter": sums["warm_it"] / tot_B,
                             "what": f"the same solve on warm iterates: control step {K_WARM + 1} of the on-device closed loop "
                                     "(shifted previous solution), device-resident inputs"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                         "traffic_source": "static: profiles/traffic.json, one `ncu --set full` capture of the same kernel, batch and "
                                           "dimensions (not measured in this run)" if traffic is not None else None,
                         "peak_source": "self-measured register-resident DFMA chains (b2mpc_dfma_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry; nominal B200 FP64 ~37-40 TFLOP/s",
                         "flops_per_launch": fl, "kernel_ms": kern_s * 1e3,
                         "hbm": {"achieved": B * cnt["bytes"] / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": B * cnt["bytes"] / kern_s / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "BatchedMpcSolver.solve_host -> b2mpc_solve_host (pinned host buffers)"},
            "closed_loop": {"value": B * world_size / (maxs["cl_ms"] * 1e-3), "unit": "env-steps/s", "ms_per_control_step": maxs["cl_ms"],
                            "mean_qp_iter": cl_it,
                            "per_iteration_vs_warm_solve": (maxs["warm_ms"] / a.steps / (sums["warm_it"] / tot_B)) / (maxs["cl_ms"] / cl_it),
                            "what": "b2mpc_closed_loop: 40 control steps of robot_ocp_problem.py:184-260 per environment, 8 per call "
                                    "(3 launches per control step), device Philox noise; per-GPU batch as the headline; "
                                    "per_iteration_vs_warm_solve = (warm solve ms per interior-point iteration) / (closed-loop ms per iteration)",
                            "config1_single_episode_ms_per_control_step": cfg1_ms if rank == 0 else None,
                            "config1_mean_qp_iter": cfg1_it if rank == 0 else None},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
        }
        line.update(extra)
        if not a.no_cpu_baseline and world_size == 1:
            cores = os.cpu_count() or 1
            sample = a.cpu_sample or 4096
            v, dt, cit = cpu_arm(a, sample, cores)
            reps = 1
            while dt * reps < 8.0 and reps < 64:       # ~10 s of CPU work in total
                v2, dt2, _ = cpu_arm(a, sample, cores)
                v = max(v, v2); reps += 1
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} problems of the same workload x {reps} repetitions (best), "
                                              "pthreads over problems; oracle/mpc_oracle.c = CPU restatement of the acados "
                                              "algorithm, not acados (parity unpinned)"}
        print(json.dumps(line), flush=True)
    if a.no_extra_configs:
        solver.close()
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
