#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched AquaGym step on N B200s, and its HBM-roofline fraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one `aqua_step` launch: every env of this rank's slab advances by one env.step()
(aqua.py:135-213) including the in-kernel auto-reset, on actions that are already resident in
HBM (a pool of pre-generated random-action tensors).  Workload: BASELINE config 2's environment
(AquaEnv-v1: discrete actions, the default 5-obstacle "example_with_obs" layout, waves, random
boat and goal) at 2^24 envs per GPU -- "1M+ concurrent envs"; 512 MiB of state per GPU, far beyond
the 126 MB L2, so that every step streams from HBM as the timing rules require.  The exact
BASELINE sizes (65,536 and 1,048,576 envs) are L2-resident and launch-latency bound; they are
reported beside the headline in `sweep` (CUDA-graph replay), not as the roofline number.
Envs shard as independent slabs (weak scaling: fixed envs per GPU); the only collective is the
all-reduce of the 8 episode counters after the timed region.

Rank 0 prints ONE JSON line.  `--impl reference` times the CPU oracle port of the reference step
(oracle/, all host threads) on a bounded sample of the same workload; the reference itself is
pure Python and cannot travel to the GPU box.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_STEP = {"discrete_i32": 65, "continuous": 69}  # DESIGN.md section 5 / SURVEY.md section 8d
METRIC = "env-steps/sec"
WORKLOAD_ENV = "AquaEnv-v1"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 24)
    ap.add_argument("--precision", default="float32", choices=["float32", "float64"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-extras", action="store_true", help="skip sweep / rollout / f64 / cpu legs")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """Polls SM clock and throttle reasons of one GPU through NVML while the bench runs."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index, period=0.005):
        self.samples = []  # (t, sm_mhz, reasons)
        self.period = period
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indices follow PCI order; honour CUDA_VISIBLE_DEVICES if it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                index = int(vis.split(",")[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        use = inside if len(inside) >= 3 else self.samples  # short region: all samples under load
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        bits = 0
        for s in use:
            bits |= s[2]
        reasons = [n for b, n in list(self.BAD.items()) + list(self.NOTE.items()) if bits & b]
        return {"sm_mhz": float(np.median([s[1] for s in use])), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(use), "samples_in_timed_region": len(inside)}


# --------------------------------------------------------------------------- reference arm (CPU)
def oracle_world(n, seed, env_id_base=0, continuous=False):
    from oracle import aqua_oracle as ao
    from aquaticgymenv_b200.vec_env import DEFAULT_OBSTACLES
    return ao.OracleVecEnv(n, continuous=continuous, obstacles=DEFAULT_OBSTACLES, seed=seed,
                           env_id_base=env_id_base)


def cpu_step_rate(n, steps, warmup, seed, threads):
    """The oracle port stepping `n` envs of the bench workload with `threads` host threads:
    per step one aqo_step_batch over pre-generated actions (auto-reset on), wall clock."""
    import ctypes as C
    from oracle import aqua_oracle as ao
    env = oracle_world(n, seed)
    env.reset()
    rs = np.random.RandomState(1234)
    pool = [rs.randint(0, 3, n).astype(np.int32) for _ in range(4)]
    reward = np.empty(n, np.float64)
    flags = np.empty(n, np.uint8)
    L = env.L

    def one(k):
        a = pool[k % len(pool)]
        L.aqo_step_batch(C.byref(env.world), C.byref(env.state), n, a.ctypes.data_as(C.POINTER(C.c_int32)),
                         None, None, reward.ctypes.data_as(C.POINTER(C.c_double)),
                         flags.ctypes.data_as(C.POINTER(C.c_uint8)), None, 1, threads)

    for k in range(warmup):
        one(k)
    t0 = time.perf_counter()
    for k in range(steps):
        one(k)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n = 1 << 20  # bounded sample: 1/16 of one GPU's slab per step
    value, dt = cpu_step_rate(n, args.steps, args.warmup, args.seed, cores)
    sample = "%d envs x %d steps of the workload (1/16 of one GPU slab per step), %d host threads" % (
        n, args.steps, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, n_local=args.envs_per_gpu),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_local):
    return {
        "workload": "%s (discrete actions, default 5-obstacle example_with_obs layout, waves, random boat+goal), "
                    "random actions resident in HBM, auto-reset in kernel" % WORKLOAD_ENV,
        "envs_per_gpu": n_local, "envs_total": n_local * args.gpus, "precision": args.precision,
        "action_dtype": "int32", "bytes_per_env_step": BYTES_PER_STEP["discrete_i32"],
        "parallelism": "env slabs x%d, no data-path collective" % args.gpus,
        "l2": "state %.0f MiB per GPU > 126 MB L2: inputs larger than L2, no flush needed" % (
            n_local * 32 / 2 ** 20) if n_local * 32 > 140e6 else "L2-resident working set",
    }


# --------------------------------------------------------------------------- GPU arm
def timed_steps(torch, env, pool, steps, dist_on):
    """K launches bracketed by barrier + synchronize; returns elapsed seconds (device, CUDA events)."""
    import torch.distributed as dist
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    start.record()
    for k in range(steps):
        env.step_flags(pool[k % len(pool)])
    end.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if dist_on:
        dist.barrier()
    return start.elapsed_time(end) * 1e-3, t0, t1


def graph_rate(torch, n, seed, precision, steps_per_graph=50, replays=6):
    """Small (L2-resident, launch-bound) slabs: `steps_per_graph` step launches captured in one
    CUDA graph, replayed; returns env-steps/s."""
    from aquaticgymenv_b200 import gym_aqua
    env = gym_aqua.make(WORKLOAD_ENV, num_envs=n, seed=seed, precision=precision)
    env.reset()
    pool = [env.random_actions(k) for k in range(4)]
    for k in range(8):
        env.step_flags(pool[k % 4])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for k in range(steps_per_graph):
            env.step_flags(pool[k % 4])
    g.replay()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(replays):
        g.replay()
    end.record()
    torch.cuda.synchronize()
    dt = start.elapsed_time(end) * 1e-3
    stats = env.stats()
    assert stats["steps"] == n * (8 + steps_per_graph * (replays + 1)), stats
    env.close()
    return n * steps_per_graph * replays / dt


def step_rate(torch, env_id, n, seed, precision, obstacles=None, steps=100):
    """Other BASELINE configs, same method as the headline (single-step launches, resident actions)."""
    from aquaticgymenv_b200 import gym_aqua
    kw = {}
    if obstacles == "dense64":
        kw["obstacles"] = np.load(os.path.join(ROOT, "tests", "golden", "dense64_table.npy"))
    env = gym_aqua.make(env_id, num_envs=n, seed=seed, precision=precision, **kw)
    env.reset()
    pool = [env.random_actions(k) for k in range(4)]
    for k in range(10):
        env.step_flags(pool[k % 4])
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for k in range(steps):
        env.step_flags(pool[k % 4])
    end.record()
    torch.cuda.synchronize()
    dt = start.elapsed_time(end) * 1e-3
    env.close()
    cont = "Continuous" in env_id
    es = 4 if precision == "float32" else 8
    bpe = (7 * es + 4) + (8 if cont else 4) + (5 * es + 4) + es + 1
    peak_gbs, _ = peaks()
    return {"env": env_id + ("+dense64" if obstacles else ""), "envs": n, "precision": precision,
            "env_steps_per_s": n * steps / dt, "bytes_per_env_step": bpe,
            "hbm_frac": bpe * n * steps / dt / 1e9 / peak_gbs}


def rollout_rate(torch, n, seed, precision, steps=100):
    from aquaticgymenv_b200 import gym_aqua
    env = gym_aqua.make(WORKLOAD_ENV, num_envs=n, seed=seed, precision=precision)
    env.reset()
    env.rollout_random(10, 0)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    env.rollout_random(steps, 10)
    end.record()
    torch.cuda.synchronize()
    env.close()
    return n * steps / (start.elapsed_time(end) * 1e-3)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from aquaticgymenv_b200 import gym_aqua, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist_on = world > 1
    if dist_on:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_local = args.envs_per_gpu
    base, _ = sharding.slab(n_local * world, rank, world)
    env = gym_aqua.make(WORKLOAD_ENV, num_envs=n_local, seed=args.seed, env_id_base=base,
                        precision=args.precision)
    env.reset()
    pool = [env.random_actions(k) for k in range(4)]  # inputs resident in HBM before timing

    sampler = ClockSampler(local_rank)
    sampler.start()
    for k in range(max(args.warmup, 3)):
        env.step_flags(pool[k % len(pool)])
    env.stats(clear=True)
    elapsed, t0, t1 = timed_steps(torch, env, pool, args.steps, dist_on)
    sampler.stop()
    clocks = sampler.summary(t0, t1)

    el = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_max = float(el.item())
    stats_t = env.stats_tensor().clone()
    sharding.all_reduce_stats(stats_t)  # the only collective of the path: 8 int64 counters
    from aquaticgymenv_b200.vec_env import stats_dict
    stats = stats_dict(stats_t.cpu().tolist())
    total_steps = n_local * world * args.steps
    assert stats["steps"] == total_steps, (stats["steps"], total_steps)
    value = total_steps / elapsed_max

    # ---- end to end through the host-buffer C-ABI call (pinned host memory both ways) ----
    e2e_value = None
    np_real = np.float32 if args.precision == "float32" else np.float64
    es = np.dtype(np_real).itemsize
    if args.e2e_steps > 0:
        h_act = torch.empty(n_local, dtype=torch.int32, pin_memory=True)
        h_act.copy_(pool[0].cpu())
        h_obs = torch.empty(5, n_local, dtype=getattr(torch, args.precision), pin_memory=True)
        h_rew = torch.empty(n_local, dtype=getattr(torch, args.precision), pin_memory=True)
        h_flg = torch.empty(n_local, dtype=torch.uint8, pin_memory=True)
        a_np, o_np, r_np, f_np = h_act.numpy(), h_obs.numpy(), h_rew.numpy(), h_flg.numpy()
        for _ in range(3):
            env.step_host(a_np, o_np, r_np, f_np)
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        te0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            env.step_host(a_np, o_np, r_np, f_np)
        torch.cuda.synchronize()
        e2e_dt = time.perf_counter() - te0
        ed = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        if dist_on:
            dist.all_reduce(ed, op=dist.ReduceOp.MAX)
        e2e_value = n_local * world * args.e2e_steps / float(ed.item())
        assert int(f_np.max()) <= 9 and np.isfinite(r_np).all()

    peak_gbs, peak_src = peaks()
    bpe = BYTES_PER_STEP["discrete_i32"]
    achieved = bpe * n_local * args.steps / elapsed / 1e9  # this rank's kernel, GB/s
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get("envs") == n_local and tj.get("precision") == args.precision:
            traffic = tj.get("dram_bytes_per_launch")

    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * elapsed_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "float32" else "f64", "data": "synthetic",
        "config": workload_config(args, n_local),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bpe * n_local,
                     # whole warps of the slab run the streaming kernel (launch_step, aqua_kernels.cu)
                     "kernel": "aqua::step_stream_kernel<%s,I32,waves>" % ("float,4" if args.precision == "float32" else "double,2")},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": 4 * n_local * world,
                "d2h_bytes_per_step": (5 * es + es + 1) * n_local * world, "steps": args.e2e_steps,
                "api": "aqua_step_host (AquaVecEnv.step_host): int32 actions in, obs[5,N]+reward+flags out, pinned"},
        "gpu_launches": args.steps,
        "clocks": clocks,
        "episode_stats": {k: stats[k] for k in ("episodes", "collided", "timeout", "success",
                                                "mean_episode_length", "mean_episode_return")},
    }

    if rank == 0 and world == 1 and not args.no_extras:
        extras = {}
        try:
            # envs/GPU sweep: 50 step launches per CUDA graph, replayed.
            # Up to ~2^20 envs the slab (41 B/env of state+outputs) is L2-resident and the step is
            # launch/latency-bound; from 2^22 it streams from HBM.
            extras["sweep_cuda_graph"] = []
            for n in (1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24):
                r = graph_rate(torch, n, args.seed, args.precision)
                extras["sweep_cuda_graph"].append({
                    "envs": n, "env_steps_per_s": r, "us_per_step": 1e6 * n / r,
                    "hbm_frac_algorithmic": r * BYTES_PER_STEP["discrete_i32"] / 1e9 / peaks()[0],
                    "regime": "launch/latency-bound, L2-resident" if n <= (1 << 20) else "HBM-streaming"})
            extras["rollout_kernel_env_steps_per_s"] = {
                str(n): rollout_rate(torch, n, args.seed, args.precision) for n in (65536, 1 << 20, n_local)}
            extras["other_configs"] = [
                step_rate(torch, "AquaContinuousEnv-v1", 1 << 24, args.seed, "float32"),
                step_rate(torch, "AquaEnv-v1", 1 << 23, args.seed, "float64"),
                step_rate(torch, "AquaEnv-v0", 1 << 22, args.seed, "float32", obstacles="dense64"),
            ]
        except Exception as e:  # extras must never cost the headline
            extras["error"] = repr(e)
        line["extras"] = extras
        cores = host_cores()
        n_cpu = 1 << 20
        probe, _ = cpu_step_rate(n_cpu, 3, 1, args.seed, cores)     # calibrate the bounded sample
        cpu_steps = int(max(5, min(4000, args.cpu_seconds * probe / n_cpu)))
        v, dt = cpu_step_rate(n_cpu, cpu_steps, 2, args.seed, cores)
        line["cpu_baseline"] = {
            "value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": "%d envs x %d steps of the same workload, C oracle port (float64), %d threads, %.1f s" % (
                n_cpu, cpu_steps, cores, dt)}
    elif rank == 0:
        line["cpu_baseline"] = None

    env.close()
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
