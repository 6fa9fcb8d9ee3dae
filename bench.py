#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched AquaGym step on N B200s, and its HBM-roofline fraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config headline|c2|c3|c4|stream] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A bench "step" = one pass of the hot path over one batch of synthetic input = ONE launch: with
aqua_step_many (the default) every env of the job advances by `--steps-per-launch` (64) env.step()
calls (aqua.py:135-213, auto-reset included) on a tape of random actions that is resident in HBM
before the timed region; with `--mode single` by one.  `value` counts env-steps (envs x env.step()
calls) per second; `ms_per_step` is per launch; `config.env_steps_per_step` says how many env-steps
one bench step is.

Workloads (BASELINE.json):
  headline  the metric's own config: AquaEnv-v1 (discrete actions, default 5-obstacle layout, waves,
            random boat + goal), 1,048,576 envs per GPU (weak scaling: N GPUs step N x 1M envs)
  c2        the same env, 65,536 envs on one GPU
  c3        AquaContinuousEnv-v1, thrust pairs, 1,048,576 envs in total sharded over the N GPUs
  c4        the 64-obstacle stress map, discrete, 4,194,304 envs in total sharded over the N GPUs
  stream    AquaEnv-v1, 2^24 envs per GPU, ONE launch per step (the HBM-streaming kernel)
aqua_step_many reads the action tape [K][N] and writes reward + flags [K][N] per env-step (tape +
outputs exceed the 126 MB L2 at the headline size); the state stays in registers between the
env-steps of a launch.
`--mode single` runs one aqua_step launch per step instead (what a policy in the loop uses); at the
L2-resident sizes it is captured in a CUDA graph.  Both figures of the headline workload are in
the line (`extras.single_step`).  Envs shard as independent slabs; the only collective is the
all-reduce of the 9 episode counters after the timed region.

Besides the contract's keys the line carries: `roofline.traffic` (DRAM bytes per launch of the committed
ncu capture of exactly this configuration, profiles/roofline_traffic.json), `roofline.issue_slots` (the
kernel is bound by instruction issue: warp instructions of that capture over the launch time and SM
clock measured here), `roofline.frac_of_bytes_moved`; `cpu_baseline.port` (the C oracle port beside the
Python reference), `cpu_baseline.c1_single_env` (BASELINE config 1: the single-env drop-in against the
reference's gym path, steps/s) and `cpu_baseline.parity_check` (the measured call against the oracle in
the same run: flag mismatches and state drift, float64 and float32); `extras` with the other BASELINE
configs at their literal sizes, the one-launch-per-step figures and the float64 (flag-exact) layout.

Rank 0 prints ONE JSON line.  `--impl reference` drives the UNMODIFIED reference env
(baseline/_ref/gym_aqua, copied there by __graft_entry__.build()) through a gym-style vector env
on all host cores (oracle/ref_vector_env.py) and never loads the product library.
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

METRIC = "env-steps/sec"
# DESIGN.md section 5 / SURVEY.md section 8d: algorithmic bytes of ONE env.step through a
# single-step launch (state in + out, action, reward, flags); the metric's roofline convention
BYTES_PER_STEP = {False: 65, True: 69}

CONFIGS = {
    "headline": dict(env="AquaEnv-v1", per_gpu=1 << 20, total=None, table=None, mode="many",
                     what="AquaEnv-v1 (discrete actions, default 5-obstacle example_with_obs layout, waves, random "
                          "boat+goal): BASELINE.json metric config, 1,048,576 envs per GPU"),
    "c2": dict(env="AquaEnv-v1", per_gpu=65536, total=None, table=None, mode="many",
               what="BASELINE config 2: AquaEnv-v1, example_with_obs layout, 65,536 envs per GPU"),
    "c3": dict(env="AquaContinuousEnv-v1", per_gpu=None, total=1 << 20, table=None, mode="many",
               what="BASELINE config 3: AquaContinuousEnv-v1, thrust pairs in [0.2,0.5]^2, 1,048,576 envs in total"),
    "c4": dict(env="AquaEnv-v0", per_gpu=None, total=1 << 22, table="dense64", mode="many",
               what="BASELINE config 4: 64-obstacle stress map (tests/golden/dense64_table.npy), discrete, "
                    "4,194,304 envs in total"),
    "stream": dict(env="AquaEnv-v1", per_gpu=1 << 24, total=None, table=None, mode="single",
                   what="AquaEnv-v1, 2^24 envs per GPU (state 512 MiB > L2), one aqua_step launch per step"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="timed launches")
    ap.add_argument("--warmup", type=int, default=5, help="untimed launches before (at least 3)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default=None, choices=["many", "single"])
    ap.add_argument("--envs-per-gpu", type=int, default=None, help="override the config's slab size")
    ap.add_argument("--steps-per-launch", type=int, default=64)
    ap.add_argument("--precision", default="float32", choices=["float32", "float64"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-extras", action="store_true", help="skip sweep / other configs / cpu legs")
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


def slab_size(cfg, world, override=None):
    if override:
        return int(override)
    if cfg["per_gpu"] is not None:
        return cfg["per_gpu"]
    return cfg["total"] // world


def workload_config(name, cfg, n_local, world, mode, spl, precision, l2_note):
    cont = "Continuous" in cfg["env"]
    return {
        "workload": cfg["what"] + "; random actions resident in HBM, auto-reset in kernel",
        "name": name, "env_id": cfg["env"], "obstacles": cfg["table"] or "env default",
        "envs_per_gpu": n_local, "envs_total": n_local * world, "precision": precision,
        "action_dtype": "float32x2" if cont else "int32",
        "mode": ("aqua_step_many: %d steps per launch, action tape in, reward+flags out per step" % spl) if mode == "many"
                else "aqua_step: one launch per step",
        "steps_per_launch": spl if mode == "many" else 1,
        "env_steps_per_step": n_local * world * (spl if mode == "many" else 1),
        "bytes_per_env_step": BYTES_PER_STEP[cont],
        "parallelism": "env slabs x%d, no data-path collective" % world,
        "l2": l2_note,
    }


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """Polls SM clock and throttle reasons of one GPU through NVML while the bench runs."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index, period=0.002):
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

    def summary(self, t0, t1, t2):
        """[t0, t1] = the timed region, [t1, t2] = the same launches repeated right behind it for long
        enough to be sampled (the timed region of a short run lasts less than a sampling period)."""
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        load = [s for s in self.samples if t0 <= s[0] <= t2]
        use = load if load else self.samples
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        bits = 0
        for s in use:
            bits |= s[2]
        reasons = [n for b, n in list(self.BAD.items()) + list(self.NOTE.items()) if bits & b]
        return {"sm_mhz": float(np.median([s[1] for s in use])), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(use), "samples_in_timed_region": len(inside),
                "window": "timed region + %.2f s of the same launches behind it" % (t2 - t1)}


def bind_to_gpu_cores(local_rank):
    """Pin this rank to the host cores next to its GPU (NVML affinity) before it allocates pinned
    memory: first touch then places the staging buffers on the GPU's NUMA node.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cores = {64 * w + b for w, bits in enumerate(words) for b in range(64) if (int(bits) >> b) & 1}
        cores &= set(os.sched_getaffinity(0))
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------- CPU legs
def reference_rate(cfg, n_envs, steps, warmup, cores, seed):
    """The unmodified reference env of `cfg` through the vector env on `cores` processes."""
    from oracle import ref_vector_env as rv
    kw = {}
    if cfg["table"] == "dense64":
        tab = np.load(os.path.join(ROOT, "tests", "golden", "dense64_table.npy"))
        kw["obstacles"] = [(np.array([r[0], r[1]]), "c", r[3]) if r[2] == 0.0 else (np.array([r[0], r[1]]), "r", (r[3], r[4]))
                           for r in tab]
    return rv.time_reference(cfg["env"], n_envs, cores, steps, warmup, seed=seed,
                             continuous="Continuous" in cfg["env"], **kw)


def port_rate(n, steps, warmup, seed, threads):
    """The C oracle port (oracle/aqua_oracle.c, float64) stepping `n` AquaEnv-v1 envs with `threads`
    host threads: a second CPU figure (compiled code instead of the reference's interpreter)."""
    import ctypes as C
    from oracle import aqua_oracle as ao
    tab = np.load(os.path.join(ROOT, "tests", "golden", "v1_table.npy"))
    env = ao.OracleVecEnv(n, continuous=False, obstacles=tab, seed=seed)
    env.reset()
    rs = np.random.RandomState(1234)
    pool = [rs.randint(0, 3, n).astype(np.int32) for _ in range(4)]
    reward = np.empty(n, np.float64)
    flags = np.empty(n, np.uint8)

    def one(k):
        a = pool[k % len(pool)]
        env.L.aqo_step_batch(C.byref(env.world), C.byref(env.state), n, a.ctypes.data_as(C.POINTER(C.c_int32)),
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
    """`--impl reference`: rank 0 times the reference's own CPU implementation of the path on the
    box's host cores, on this arm's config / metric / unit; the other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = host_cores()
    cfg = CONFIGS[args.config]
    world = max(args.gpus, 1)
    n_local = slab_size(cfg, world, args.envs_per_gpu)
    mode = args.mode or cfg["mode"]
    from oracle import ref_vector_env as rv
    # bounded sample: 4096 envs per host core per step (65,536 on 16 cores), a few seconds in all
    n_ref = min(n_local * world, 4096 * cores)
    if rv.reference_root() is not None:
        value, dt, n_ref = reference_rate(cfg, n_ref, args.steps, args.warmup, cores, args.seed)
        kind = "reference"
        sample = ("%d envs x %d steps of the workload (of %d envs): unmodified reference gym_aqua env, "
                  "%d worker processes x %d envs in a lock-step vector env with auto-reset, random actions" % (
                      n_ref, args.steps, n_local * world, cores, n_ref // cores))
    else:  # the reference copy did not travel: the compiled restatement, said so
        n_ref = 1 << 20
        value, dt = port_rate(n_ref, args.steps, args.warmup, args.seed, cores)
        kind = "port"
        sample = "baseline/_ref missing: C oracle port, %d envs x %d steps, %d threads" % (n_ref, args.steps, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args.config, cfg, n_local, world, mode, args.steps_per_launch, args.precision,
                                       "n/a (host)"), envs_stepped_by_this_arm=n_ref),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def make_env(torch, cfg, n_local, base, seed, precision, **kw):
    from aquaticgymenv_b200 import gym_aqua
    if cfg["table"] == "dense64":
        kw["obstacles"] = np.load(os.path.join(ROOT, "tests", "golden", "dense64_table.npy"))
    return gym_aqua.make(cfg["env"], num_envs=n_local, seed=seed, env_id_base=base, precision=precision, **kw)


class ManyRunner(object):
    """K steps per launch: the tape of random actions and the outputs live in HBM."""

    def __init__(self, torch, env, spl):
        self.torch, self.env, self.spl = torch, env, spl
        n16 = (env.num_envs + 15) // 16 * 16
        if env.continuous:
            self.tape = torch.full((spl, n16, 2), 0.35, dtype=torch.float32, device=env.device)
        else:
            self.tape = torch.full((spl, n16), 2, dtype=torch.int32, device=env.device)
        for k in range(spl):  # the synthetic random policy (test_random.py:19-20), Philox ACTION stream
            self.tape[k, :env.num_envs] = env.random_actions(k)
        self.rew = torch.empty(spl, n16, dtype=env.dtype, device=env.device)
        self.flg = torch.empty(spl, n16, dtype=torch.uint8, device=env.device)
        self.bytes_resident = self.tape.numel() * self.tape.element_size() + self.rew.numel() * self.rew.element_size() + self.flg.numel()
        self.launches = 0

    per_launch = property(lambda self: self.spl)   # env-steps per env and launch

    def run(self, launches):
        for _ in range(launches):
            self.env.step_many(self.tape, self.rew, self.flg)
        self.launches += launches

    def l2_note(self):
        mb = self.bytes_resident / 1e6
        return ("inputs larger than L2: action tape + reward/flags outputs of one launch = %.0f MB (> 126 MB L2), plus "
                "%.0f MB of state; no flush" % (mb, self.env.num_envs * 64 / 1e6)) if mb > 126 else \
            "L2-resident working set (%.0f MB tape + outputs per launch): launch/issue-bound regime, no flush" % mb


class SingleRunner(object):
    """One aqua_step launch per step over a pool of resident action tensors; L2-resident slabs replay
    a CUDA graph of `chunk` launches."""

    def __init__(self, torch, env, graph_chunk=0):
        self.torch, self.env = torch, env
        self.pool = [env.random_actions(k) for k in range(4)]
        self.launches = 0
        self.graph, self.chunk = None, graph_chunk
        if graph_chunk:
            for k in range(4):
                env.step_flags(self.pool[k])
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                for k in range(graph_chunk):
                    env.step_flags(self.pool[k % 4])

    per_launch = 1

    def run(self, steps):
        done = 0
        while self.graph is not None and steps - done >= self.chunk:
            self.graph.replay()
            done += self.chunk
            self.launches += self.chunk
        for k in range(steps - done):
            self.env.step_flags(self.pool[k % 4])
            self.launches += 1

    def l2_note(self):
        mb = self.env.num_envs * 65 / 1e6
        return ("inputs larger than L2: %.0f MB move per step (state %.0f MiB); no flush" % (mb, self.env.num_envs * 32 / 2 ** 20)
                if mb > 126 else "L2-resident working set (%.0f MB per step)%s: launch-latency-bound regime, no flush" % (
                    mb, ", CUDA graph of %d launches" % self.chunk if self.graph is not None else ""))


def timed(torch, runner, steps, dist_on):
    """Exactly `steps` steps bracketed by barrier + synchronize; returns (device seconds by CUDA
    events on the launching stream, host t0, t1, launches)."""
    import torch.distributed as dist
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = runner.launches
    t0 = time.perf_counter()
    start.record()
    runner.run(steps)
    end.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if dist_on:
        dist.barrier()
    return start.elapsed_time(end) * 1e-3, t0, t1, runner.launches - l0


def measure(torch, name, world, rank, args, mode=None, n_local=None, steps=None, warmup=None, precision=None,
            dist_on=False, sustain=0.0, spl=None):
    """One workload on this rank's slab -> dict (rates are this rank's; the caller reduces)."""
    from aquaticgymenv_b200 import sharding
    cfg = CONFIGS[name]
    mode = mode or cfg["mode"]
    precision = precision or args.precision
    spl = spl or args.steps_per_launch
    n_local = n_local or slab_size(cfg, world)
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    base, _ = sharding.slab(n_local * world, rank, world)
    env = make_env(torch, cfg, n_local, base, args.seed, precision)
    env.reset()
    if mode == "many":
        runner = ManyRunner(torch, env, spl)
    else:
        runner = SingleRunner(torch, env, graph_chunk=32 if n_local * 65 < 126e6 else 0)
    # warm-up: W launches, at least three (stash allocation, lazy initialisation, clocks)
    runner.run(max(warmup, 3))
    torch.cuda.synchronize()
    env.stats(clear=True)
    elapsed, t0, t1, launches = timed(torch, runner, steps, dist_on)
    stats_t = env.stats_tensor().clone()
    t2 = t1
    if sustain > 0:  # the same launches for long enough that the clock sampler sees the GPU under this load
        while time.perf_counter() - t1 < sustain:
            runner.run(steps)
            torch.cuda.synchronize()
        t2 = time.perf_counter()
    return {"name": name, "cfg": cfg, "mode": mode, "n_local": n_local, "steps": steps, "elapsed": elapsed,
            "launches": launches, "t": (t0, t1, t2), "stats_t": stats_t, "continuous": env.continuous,
            "precision": precision, "l2": runner.l2_note(), "spl": spl, "env": env, "runner": runner,
            "per_launch": runner.per_launch}


def summarise(torch, m, world, dist_on):
    """Reduce one measurement over the ranks (max time, summed counters) -> JSON-able dict."""
    import torch.distributed as dist
    from aquaticgymenv_b200 import sharding
    from aquaticgymenv_b200.vec_env import stats_dict
    el = torch.tensor([m["elapsed"]], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_max = float(el.item())
    sharding.all_reduce_stats(m["stats_t"])  # the only collective of the path: 9 int64 counters
    stats = stats_dict(m["stats_t"].cpu().tolist())
    total = m["n_local"] * world * m["steps"] * m["per_launch"]
    assert stats["steps"] == total, (stats["steps"], total)  # the kernels counted exactly the timed env-steps
    peak_gbs, peak_src = peaks()
    es = 8 if m["precision"] == "float64" else 4
    # single-step launch: 7 state words + time in, action in, 5 state words + time out, reward, flags
    bpe = (7 * es + 4) + (8 if m["continuous"] else 4) + (5 * es + 4) + es + 1
    achieved = bpe * m["n_local"] * m["steps"] * m["per_launch"] / m["elapsed"] / 1e9  # this rank's kernel, GB/s
    moved = None
    if m["mode"] == "many":
        # what one env-step really moves: action in, reward + flags out; per launch the state planes
        # in + out (2 x 8 words), the four ring tags (16 B) and the record of the env's next episode
        # (32 B); per finished episode (episodes / env-steps of the timed region) a record copied in
        # (32 B) and one drawn again (32 B + its tag).  The measured DRAM bytes of the same launch are
        # `traffic` (ncu), when this exact configuration was captured.
        per_launch = 2 * (7 * es + 4) + 16 + 32
        per_episode = 32 + 36
        moved = ((8 if m["continuous"] else 4) + es + 1 + per_launch / m["spl"] +
                 per_episode * stats["episodes"] / max(stats["steps"], 1))
    return {
        "value": total / elapsed_max, "elapsed_max_s": elapsed_max, "ms_per_step": 1e3 * elapsed_max / m["steps"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                     "traffic": traffic_of(m), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bpe * m["n_local"] * m["per_launch"],
                     "bytes_per_env_step": bpe,
                     "bytes_per_env_step_moved_by_this_kernel": moved,
                     "frac_of_bytes_moved": None if moved is None else achieved / peak_gbs * moved / bpe,
                     "note": ("aqua::step_many_kernel keeps the state in registers for the steps of a launch: it is bound by "
                              "instruction issue, not HBM; `frac` is the metric's convention (env-steps/s x %d B / peak), "
                              "`frac_of_bytes_moved` what the kernel itself moves" % bpe) if m["mode"] == "many" else
                             "aqua::step_stream_kernel / step_narrow_kernel: one launch per step, %d B per env-step" % bpe,
                     "kernel": "aqua::step_many_kernel" if m["mode"] == "many" else
                               ("aqua::step_stream_kernel" if m["n_local"] >= 681984 else "aqua::step_narrow_kernel")},
        "gpu_launches": m["launches"],
        "episode_stats": {k: stats[k] for k in ("episodes", "collided", "timeout", "success", "forced_resets",
                                                "mean_episode_length", "mean_episode_return")},
    }


def traffic_entry(m):
    """The committed ncu capture of exactly this (config, envs, mode, steps per launch, precision), or None."""
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(tp):
        return None
    with open(tp) as f:
        tj = json.load(f)
    key = "%s:%d:%s:%d:%s" % (m["name"], m["n_local"], m["mode"], m["spl"] if m["mode"] == "many" else 1, m["precision"])
    return tj.get("captures", {}).get(key)


def traffic_of(m):
    """DRAM bytes per launch of this workload's kernel from that capture; else null."""
    ent = traffic_entry(m)
    return None if ent is None else ent.get("dram_bytes_per_launch")


def brief(s, m, world):
    return {"config": m["name"], "mode": m["mode"], "envs_per_gpu": m["n_local"], "envs_total": m["n_local"] * world,
            "steps_per_launch": m["spl"] if m["mode"] == "many" else 1, "precision": m["precision"],
            "env_steps_per_s": s["value"], "us_per_env_step_of_all_envs": 1e3 * s["ms_per_step"] / m["per_launch"],
            "hbm_frac": s["roofline"]["frac"], "bytes_per_env_step": s["roofline"]["bytes_per_env_step"], "l2": m["l2"]}


def e2e_leg(torch, env, args, world, dist_on):
    """The host-buffer C-ABI call (aqua_step_host): pinned uint8 actions in, obs + reward + flags out,
    copies inside the timed region.  Also a plain pinned device-to-host copy of the same bytes, as the
    ceiling of the link on this box."""
    import torch.distributed as dist
    n = env.num_envs
    tdt = torch.float32 if env.precision == 0 else torch.float64
    es = 4 if env.precision == 0 else 8
    h_act = torch.empty(n, dtype=torch.uint8, pin_memory=True)   # the action index fits a byte (AQUA_ACT_U8)
    h_act.copy_(env.random_actions(0).cpu().to(torch.uint8))
    h_obs = torch.empty(5, n, dtype=tdt, pin_memory=True)
    h_rew = torch.empty(n, dtype=tdt, pin_memory=True)
    h_flg = torch.empty(n, dtype=torch.uint8, pin_memory=True)
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
    assert int(f_np.max()) <= 9 and np.isfinite(r_np).all()
    d2h = env.host_d2h_bytes_per_env() * n
    # ceiling: the same number of bytes, device -> pinned host, one plain copy per step
    src = torch.empty(d2h, dtype=torch.uint8, device=env.device)
    dst = torch.empty(d2h, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    tc0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    copy_dt = time.perf_counter() - tc0
    cd = torch.tensor([copy_dt], dtype=torch.float64, device="cuda")
    if dist_on:
        dist.all_reduce(cd, op=dist.ReduceOp.MAX)
    return {"value": n * world * args.e2e_steps / float(ed.item()), "unit": "env-steps/s",
            "h2d_bytes_per_step": 1 * n * world, "d2h_bytes_per_step": d2h * world, "steps": args.e2e_steps,
            "d2h_gbs_per_gpu": d2h * args.e2e_steps / float(ed.item()) / 1e9,
            "plain_pinned_d2h_copy_gbs_per_gpu": d2h * args.e2e_steps / float(cd.item()) / 1e9,
            "api": "aqua_step_host (AquaVecEnv.step_host): uint8 actions in; x, y, angle + reward + flags out every step, "
                   "goals only for the envs whose episode changed (written straight into the pinned planes); "
                   "pinned host buffers"}


def run_b200(args):
    import torch
    import torch.distributed as dist

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
    bound_cores = bind_to_gpu_cores(local_rank) if dist_on else None
    if dist_on:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    name = args.config
    cfg = CONFIGS[name]
    mode = args.mode or cfg["mode"]
    n_local = slab_size(cfg, world, args.envs_per_gpu)

    sampler = ClockSampler(local_rank)
    sampler.start()
    m = measure(torch, name, world, rank, args, mode=mode, n_local=n_local, dist_on=dist_on, sustain=0.35)
    sampler.stop()
    clocks = sampler.summary(*m["t"])
    s = summarise(torch, m, world, dist_on)

    e2e = {"value": None, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if args.e2e_steps > 0 and not m["continuous"]:
        e2e = e2e_leg(torch, m["env"], args, world, dist_on)
        e2e["host_cores_bound_to_gpu"] = bound_cores
    m["env"].close()

    # second view of the dominant kernel (it is bound by instruction issue, not by HBM): the warp
    # instructions per launch of the committed ncu capture of exactly this configuration, over the
    # launch time measured here and the SM clock sampled here -> fraction of the issue slots
    try:
        cap = traffic_entry(m)
        if cap is not None and cap.get("warp_instructions_per_launch") and clocks.get("sm_mhz"):
            props = torch.cuda.get_device_properties(local_rank)
            slots = props.multi_processor_count * 4 * clocks["sm_mhz"] * 1e6
            per_launch_s = m["elapsed"] / m["steps"]
            s["roofline"]["issue_slots"] = {
                "warp_instructions_per_launch": cap["warp_instructions_per_launch"],
                "frac": cap["warp_instructions_per_launch"] / per_launch_s / slots,
                "peak_warp_instructions_per_s": slots,
                "source": "%s; %s" % (cap.get("source"), cap.get("warp_instructions_note", ""))}
    except Exception as e:  # a supplementary figure must never cost the line
        s["roofline"]["issue_slots"] = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": s["value"], "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == "float32" else "f64", "data": "synthetic",
        "config": workload_config(name, cfg, n_local, world, mode, args.steps_per_launch, args.precision, m["l2"]),
        "roofline": s["roofline"], "e2e": e2e, "gpu_launches": s["gpu_launches"], "clocks": clocks,
        "episode_stats": s["episode_stats"],
    }

    if not args.no_extras:
        extras = {}
        try:
            def one(nm, **kw):
                many = (kw.get("mode") or CONFIGS[nm]["mode"]) == "many"
                mm = measure(torch, nm, world, rank, args, dist_on=dist_on, steps=6 if many else 128,
                             warmup=3 if many else 32, **kw)
                ss = summarise(torch, mm, world, dist_on)
                mm["env"].close()
                return brief(ss, mm, world)

            # the same workload, one launch per step (what a policy in the loop uses)
            extras["single_step"] = one(name, mode="single", n_local=n_local)
            # the other BASELINE configs at their literal sizes on this many GPUs
            extras["configs"] = [one(nm) for nm in ("c2", "c3", "c4") if nm != name]
            if world == 1 and rank == 0:
                extras["configs"] += [one("c3", mode="single"), one("c4", mode="single"), one("c2", mode="single")]
                extras["stream_2p24_single_step"] = one("stream")
                extras["float64_flag_exact"] = [one("headline", precision="float64"),
                                                one("stream", precision="float64", n_local=1 << 23)]
                # envs/GPU sweep, both launch forms
                extras["sweep"] = []
                for n in (1 << 16, 1 << 17, 1 << 19, 1 << 20, 1 << 22, 1 << 24):
                    extras["sweep"].append(one("headline", n_local=n))
                    extras["sweep"].append(one("headline", n_local=n, mode="single"))
        except Exception as e:  # extras must never cost the headline
            extras["error"] = repr(e)
        line["extras"] = extras
        if rank == 0 and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, cfg, n_local)
            line["cpu_baseline"]["c1_single_env"] = c1_single_env(args)  # BASELINE config 1, both sides
    elif rank == 0:
        line["cpu_baseline"] = None

    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def c1_single_env(args, steps=10000):
    """BASELINE config 1: AquaEnv-v0 (discrete, no obstacles), ONE env, random policy, `steps` steps,
    reset on done -- the reference's own gym path (unmodified Python env, one process) and the
    drop-in single-env class (same API; every step is one call into the library: action up, kernel,
    observation + reward + flags down, one synchronisation)."""
    out = {"steps": steps, "env": "AquaEnv-v0 (no obstacles, waves, random boat + goal), single env, action_space.sample()"}
    try:
        from aquaticgymenv_b200 import gym_aqua
        env = gym_aqua.make("AquaEnv-v0", seed=args.seed)
        env.action_space.seed(0)
        env.reset()
        for _ in range(200):
            if env.step(env.action_space.sample())[2]:
                env.reset()
        t0 = time.perf_counter()
        for _ in range(steps):
            if env.step(env.action_space.sample())[2]:
                env.reset()
        out["dropin_steps_per_s"] = steps / (time.perf_counter() - t0)
        env.close()
    except Exception as e:
        out["dropin_error"] = repr(e)
    try:
        from oracle import ref_vector_env as rv
        gym = rv.import_reference()
        np.random.seed(0)
        env = gym.make("AquaEnv-v0")
        env.action_space.seed(0)
        env.observation_space.seed(0)
        env.reset()
        t0 = time.perf_counter()
        for _ in range(steps):
            if env.step(env.action_space.sample())[2]:
                env.reset()
        out["reference_steps_per_s"] = steps / (time.perf_counter() - t0)
    except Exception as e:
        out["reference_error"] = repr(e)
    return out


def cpu_baseline(args, cfg, n_local):
    """The reference's CPU gym step on this box's host cores, same run: the unmodified Python env
    through the vector env (kind "reference"), bounded to about --cpu-seconds; the C port beside it."""
    cores = host_cores()
    out = {}
    try:
        from oracle import ref_vector_env as rv
        if rv.reference_root() is None:
            raise ImportError("baseline/_ref missing")
        n_ref = min(n_local, 2048 * cores)
        probe, _, n_ref = reference_rate(cfg, n_ref, 2, 1, cores, args.seed)
        steps = int(max(3, min(400, args.cpu_seconds * probe / n_ref)))
        v, dt, n_ref = reference_rate(cfg, n_ref, steps, 1, cores, args.seed)
        out = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "reference",
               "sample": "%d envs x %d steps of the same workload (of %d envs): unmodified reference gym_aqua env (Python), "
                         "%d worker processes in a lock-step vector env with auto-reset, %.1f s" % (
                             n_ref, steps, n_local, cores, dt)}
    except Exception as e:
        out = {"value": None, "unit": "env-steps/s", "cores": cores, "kind": "reference", "sample": "failed: %r" % (e,)}
    try:
        n_cpu = 1 << 20
        probe, _ = port_rate(n_cpu, 3, 1, args.seed, cores)
        cpu_steps = int(max(5, min(4000, 0.5 * args.cpu_seconds * probe / n_cpu)))
        v, dt = port_rate(n_cpu, cpu_steps, 2, args.seed, cores)
        out["port"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                       "sample": "%d envs x %d steps, C oracle port of the reference step (float64), %d threads, %.1f s" % (
                           n_cpu, cpu_steps, cores, dt)}
        if out.get("value") is None:
            out.update({k: out["port"][k] for k in ("value", "kind", "sample")})
    except Exception as e:
        out["port"] = {"error": repr(e)}
    try:
        out["parity_check"] = parity_check(args)
    except Exception as e:
        out["parity_check"] = {"error": repr(e)}
    return out


def parity_check(args, n=8192, launches=2, spl=64):
    """The measured call (aqua_step_many, the headline config's env) against the oracle in the same run,
    on the same random actions: the float64 instantiation flag for flag over n x launches x spl
    env-steps, and the float32 throughput layout over its first launch -- how many envs keep the
    oracle's flag history, and the drift of their state (scaled as tests/helpers.py: 100 px, 2 pi,
    0.05, 10).  The oracle is used here as the checker only."""
    import torch
    from aquaticgymenv_b200 import gym_aqua
    from oracle import aqua_oracle as ao
    res = {"envs": n, "env_steps": n * launches * spl}
    scale = np.array([100.0, 100.0, 2 * np.pi, 0.05, 0.05, 100.0, 100.0])[:, None]
    for precision in ("float64", "float32"):
        env = gym_aqua.make("AquaEnv-v1", num_envs=n, seed=args.seed + 99, precision=precision)
        ref = ao.OracleVecEnv(n, continuous=False, obstacles=env.table, seed=args.seed + 99)
        env.reset()
        ref.reset()
        same = np.ones(n, bool)
        mism = 0
        for l in range(launches if precision == "float64" else 1):
            tape = torch.stack([env.random_actions(l * spl + k) for k in range(spl)])
            n16 = (n + 15) // 16 * 16
            t16 = torch.full((spl, n16), 2, dtype=torch.int32, device=tape.device)
            t16[:, :n] = tape
            _, flg = env.step_many(t16)
            flg = flg[:, :n].cpu().numpy()
            acts = tape.cpu().numpy()
            for k in range(spl):
                _, _, f_ref, _ = ref.step(acts[k], autoreset=True)
                d = flg[k] != f_ref
                mism += int(d.sum())
                same &= ~d
        diff = env.get_state().cpu().numpy()[:7] - ref.get_state()[:7]
        diff[2] = (diff[2] + np.pi) % (2 * np.pi) - np.pi   # headings on either side of the +-pi seam are close
        err = np.abs(diff) / scale
        key = "f64" if precision == "float64" else "f32"
        res[key + "_flag_mismatches"] = mism
        res[key + "_envs_with_the_oracles_flag_history"] = int(same.sum())
        res[key + "_max_scaled_state_error_of_those"] = float(err[:, same].max())
        env.close()
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
