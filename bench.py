#!/usr/bin/env python
"""
bench.py -- env steps/sec of the VecGame hot path on N B200s (one JSON line, see the driver contract).

Workload (BASELINE.json configs[2]/[3]): M = 2^24 games per GPU, random-valid-action rollout,
one "step" = prepare() (auto-reset of finished games) + step() (move, reward/score, spawn, mask,
terminal flag) with the fp32 one-hot observation fused, replay (bit-exact) spawn tables, boards in
steady state after a 256-step burn-in.  Games are sharded over the ranks by global slot; the data
path has no collective, episode statistics are all-reduced once at the end (weak scaling).

  value      whole-job env-steps/s, device-timed (CUDA events), state resident in HBM
  e2e        same metric through the reference-facing API with HOST buffers: actions H2D from pinned
             memory, board/mask/reward/terminated D2H every step
  roofline   step kernel: algorithmic bytes (59 + 1024 = 1083 B per env step) over its
             CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference
             the CPU port of the reference (oracle/, C + OpenMP, all host threads) on a bounded
             sample of the same workload
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
BYTES_CORE = 59          # 25 B read + 34 B written per env step
BYTES_ONEHOT_F32 = 1024  # fp32 (16 classes x 16 cells) observation written per env step
BURN_IN = 256            # steps from reset() to steady-state boards


def parse_args() -> argparse.Namespace:
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", choices=["b200", "reference"], default="b200")
    p.add_argument("--games-per-gpu", type=int, default=1 << 24)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (core-only, small M)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--cpu-games", type=int, default=1 << 22, help="games per CPU step (bounded sample)")
    p.add_argument("--burn-in", type=int, default=BURN_IN, help="untimed steps from reset() to steady state")
    p.add_argument("--clock-interval-ms", type=float, default=50.0, help="NVML sampling period during the timed region")
    return p.parse_args()


def workload_name(games_per_gpu: int, burn_in: int = BURN_IN) -> str:
    return (f"random-valid-action rollout, M={games_per_gpu} games/GPU (2^{games_per_gpu.bit_length() - 1}), "
            "prepare(auto-reset)+step+spawn+mask+terminal fused with fp32 one-hot obs, replay (bit-exact) tables, "
            f"steady state after {burn_in}-step burn-in")


# -------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) -- also the --impl reference arm
# -------------------------------------------------------------------------------------------------


def time_cpu_port(games: int, steps: int, warmup: int, seed: int, budget_s: float | None = None) -> dict:
    """prepare()+step() of the CPU port on all host threads; actions (the policy) are generated outside
    the timed spans (BASELINE.md section 3 protocol).  Returns steps/s and what was run."""
    from oracle import oracle as orc

    lib = orc.load_lib()
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it for the CPU arm)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    lib.orc_set_num_threads(avail)
    cores = int(lib.orc_num_threads())
    env = orc.OracleVecGame(games, "normal")
    env.reset(seed)
    import numpy as np

    acts = np.empty((games,), dtype=np.int64)
    # a short burn-in so that resets occur; the full 256-step burn-in would dominate the CPU budget
    for t in range(warmup):
        env.prepare()
        env.random_valid_actions(1000 + t, acts)
        env.step(acts)
    spent = 0.0
    done = 0
    t_start = time.perf_counter()
    for t in range(steps):
        t0 = time.perf_counter()
        env.prepare()
        t1 = time.perf_counter()
        env.random_valid_actions(t, acts)  # policy: untimed
        t2 = time.perf_counter()
        env.step(acts)
        t3 = time.perf_counter()
        spent += (t1 - t0) + (t3 - t2)
        done += 1
        if budget_s is not None and time.perf_counter() - t_start > budget_s:
            break
    return {
        "value": games * done / spent,
        "ms_per_step": 1e3 * spent / done,
        "steps": done,
        "cores": cores,
        "games": games,
        "sample": f"CPU port of the reference (oracle/vecgame_oracle.c, OpenMP, {cores} threads): prepare()+step() on "
                  f"M={games} games x {done} steps after {warmup} warm-up steps from reset(seed={seed}), random-valid "
                  "actions generated outside the timed spans",
    }


def run_reference_arm(args: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 runs the CPU arm
    res = time_cpu_port(args.cpu_games, args.steps, max(args.warmup, 3), args.seed)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": res["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": res["steps"],
        "warmup": max(args.warmup, 3),
        "ms_per_step": res["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": workload_name(args.games_per_gpu), "games_per_step_sampled": res["games"]},
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU periodically through NVML while running."""

    REASONS = {
        0x8: "hw_slowdown",
        0x40: "hw_thermal_slowdown",
        0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int, interval_s: float = 0.05):
        super().__init__(daemon=True)
        self.index = index
        self.interval_s = interval_s
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = None
            try:
                import torch

                pr = torch.cuda.get_device_properties(index)
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self._h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self) -> None:
        if not self.ok:
            return
        nv = self._nv
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.interval_s)

    def stop(self) -> dict:
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = sorted(self.samples)[len(self.samples) // 2] if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------
# the B200 arm
# -------------------------------------------------------------------------------------------------


def load_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_ncu_traffic(games: int):
    """dram bytes per launch of the fused step kernel from the committed ncu summary, if it matches M."""
    path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        per_game = d.get("dram_bytes_per_game_fused_f32")
        return None if per_game is None else float(per_game) * games
    except Exception:
        return None


def run_b200_arm(args: argparse.Namespace) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback in ml2048_b200)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single process: not bound"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import ml2048_b200
    from ml2048_b200.sharding import reduce_episode_stats
    from ml2048_b200.vecgame import stats_to_dict

    m = args.games_per_gpu
    k_steps, warm = args.steps, max(args.warmup, 3)
    slot_base = rank * m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: `value` --------------------------------------------------------
    env = ml2048_b200.VecGame(m, ml2048_b200.reward_fn_normal, output="torch", onehot="f32", track_merged=False,
                              slot_base=slot_base, sync_free=True, device=dev)
    env.reset(args.seed)
    for _ in range(args.burn_in):
        env.prepare()
        env.step_random()
    torch.cuda.synchronize()
    snapshot = None if args.no_e2e else env.state_dict()

    STATS_EVERY = 64  # BASELINE configs[3]: episode statistics are all-reduced every 64 steps (24 integers)
    pending = []

    def run_steps(e, n: int, kernel_events=None):
        for i in range(n):
            e.prepare()
            if kernel_events is not None:
                kernel_events[i][0].record()
            e.step_random()
            if kernel_events is not None:
                kernel_events[i][1].record()
            if kernel_events is not None and (i + 1) % STATS_EVERY == 0:
                # the job's only collective: SUM of the max-tile histogram + episode/score/step sums, MAX of the
                # best score; asynchronous on NCCL's stream, nothing waits for it inside the timed loop
                snap = e.episode_stats_tensor()
                if world > 1:
                    sums, mx = snap[:23].clone(), snap[23:].clone()
                    pending.append((dist.all_reduce(sums, op=dist.ReduceOp.SUM, async_op=True),
                                    dist.all_reduce(mx, op=dist.ReduceOp.MAX, async_op=True), sums, mx))

    run_steps(env, warm)
    # first use of a torch op / an NCCL communicator loads kernels and connects ranks (tens to hundreds of ms):
    # do both once before the timed region, which repeats them every STATS_EVERY steps
    warm_stats = env.episode_stats_tensor()
    if world > 1:
        dist.all_reduce(warm_stats[:23].clone(), op=dist.ReduceOp.SUM)
        dist.all_reduce(warm_stats[23:].clone(), op=dist.ReduceOp.MAX)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank, args.clock_interval_ms * 1e-3)
    barrier()
    sampler.start()
    start.record()
    run_steps(env, k_steps, kev)
    stop.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = max_over_ranks(start.elapsed_time(stop))
    value = world * m * k_steps / (elapsed_ms * 1e-3)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / k_steps
    kernel_ms = max_over_ranks(kernel_ms)
    gaps = sorted(kev[i][1].elapsed_time(kev[i + 1][0]) for i in range(k_steps - 1)) or [0.0]
    prepare_ms = sum(gaps) / len(gaps)
    if os.environ.get("ML2048_BENCH_DEBUG"):
        print(f"[debug] prepare gaps ms: min {gaps[0]:.3f} p50 {gaps[len(gaps)//2]:.3f} p90 {gaps[int(len(gaps)*0.9)]:.3f} max {gaps[-1]:.3f}",
              file=sys.stderr)

    peak, peak_src = load_peaks()
    bytes_per_launch = (BYTES_CORE + BYTES_ONEHOT_F32) * m
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": "step_kernel<replay, onehot f32> (move+reward+spawn+mask+terminal+one-hot)",
        "achieved": achieved,
        "peak": peak,
        "peak_source": peak_src,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": load_ncu_traffic(m),
        "bytes_per_env_step": BYTES_CORE + BYTES_ONEHOT_F32,
        "kernel_ms": kernel_ms,
        "kernel_share_of_step": kernel_ms / (elapsed_ms / k_steps),
        "prepare_ms": prepare_ms,
    }

    for w0, w1, _, _ in pending:
        w0.wait()
        w1.wait()
    # final statistics: the only collective of the job (24 integers)
    stats = reduce_episode_stats(env.episode_stats_tensor())
    stats_d = stats_to_dict(stats)

    extras = {}
    if not args.no_extras and rank == 0:
        extras = measure_extras(ml2048_b200, torch, dev, args.seed)

    # ---- end-to-end arm: host buffers through the reference-facing API -----------------------
    e2e = None
    if not args.no_e2e:
        # record the (valid, random) actions of the next warm+k steps of this exact trajectory, then
        # rewind the environment and replay them from pinned HOST memory through step(actions)
        env.load_state_dict(snapshot)
        total = warm + k_steps
        host_actions = torch.empty((total, m), dtype=torch.uint8, pin_memory=True)
        for t in range(total):
            env.prepare()
            env.step_random(return_actions=True)
            host_actions[t].copy_(env._actions_out, non_blocking=True)
        torch.cuda.synchronize()
        env.load_state_dict(snapshot)
        env.configure(output="numpy", sync_free=False)
        consumed = 0

        def e2e_step(t: int) -> int:
            (idx,) = env.prepare()                # D2H: reset count + indices
            keys = ("state", "valid_actions", "reward", "terminated")  # D2H: what a rollout consumer reads
            res = env.step(host_actions[t], fetch=keys)  # H2D: M action bytes
            got = 0
            for key in keys:
                got += res[key].nbytes
            return got + idx.nbytes + 8

        for t in range(warm):
            e2e_step(t)
        barrier()
        t0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        d2h = 0
        for t in range(warm, total):
            d2h += e2e_step(t)
        ev1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), wall_ms))
        # the replayed host actions were recorded from this very trajectory, so every one of them must be a valid move
        invalid_last = int(env._invalid.sum().item())
        # second regime: the policy lives on the GPU, the host only reads reward + terminated (5 B per game)
        env.load_state_dict(snapshot)  # rewind again: the recorded actions belong to this trajectory
        del snapshot
        light_steps = min(total - 3, 20)
        for t in range(3):
            env.prepare()
            env.step(host_actions[t], fetch=("reward", "terminated"))
        barrier()
        t0 = time.perf_counter()
        for t in range(3, 3 + light_steps):
            env.prepare()
            res = env.step(host_actions[t], fetch=("reward", "terminated"))
        torch.cuda.synchronize()
        light_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / light_steps
        invalid_last = max(invalid_last, int(env._invalid.sum().item()))
        e2e = {
            "value": world * m * k_steps / (e2e_ms * 1e-3),
            "unit": UNIT,
            "h2d_bytes_per_step": int(m * world),
            "d2h_bytes_per_step": int(d2h / k_steps * world),
            "ms_per_step": e2e_ms / k_steps,
            "invalid_moves_in_last_step": invalid_last,
            "light": {"value": world * m / (light_ms * 1e-3), "ms_per_step": light_ms,
                      "d2h": "reward + terminated only (5 B per game); observations stay on the device"},
            "api": "VecGame.prepare() -> (indices,); VecGame.step(uint8 actions in pinned host memory) -> "
                   "state, valid_actions, reward, terminated as host arrays",
        }

    cpu_baseline = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        res = time_cpu_port(args.cpu_games, 60, 3, args.seed, budget_s=20.0)
        cpu_baseline = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": k_steps,
            "warmup": warm,
            "ms_per_step": elapsed_ms / k_steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "u8",
            "data": "synthetic",
            "config": {
                "workload": workload_name(m, args.burn_in),
                "games_per_gpu": m,
                "global_games": m * world,
                "sharding": f"dp{world}: contiguous global slots, no data-path collective; statistics (24 ints) "
                            f"all-reduced asynchronously every {STATS_EVERY} steps",
                "l2": "inputs larger than L2 (boards 268 MB, one-hot 17 GB per GPU)",
                "rng": "replay tables (bit-exact mode); actions: uniform over valid, Philox, in-kernel",
                "host_affinity": numa,
            },
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": (4 if os.environ.get("ML2048_PREPARE", "").startswith("s") else 2) * k_steps,
            "launches_per_step": ("prepare_count, prepare_scan, prepare_apply, step_kernel" if os.environ.get("ML2048_PREPARE", "").startswith("s")
                                  else "prepare_fused_kernel (one cooperative launch), step_kernel"),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "episode_stats": {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in stats_d.items()},
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_extras(ml2048_b200, torch, dev, seed: int) -> dict:
    """Secondary lines (not the headline): core-only path, Philox mode, narrower one-hot types, and the
    training shapes of run_train3.py (BASELINE configs[1]) replayed as a 16-step CUDA graph."""
    out = {}

    def timed(env, n, graph_steps=0):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if graph_steps:
            roll = ml2048_b200.GraphedRollout(env, graph_steps, window=graph_steps * 16)
            roll.replay(1)
            torch.cuda.synchronize()
            a.record()
            roll.replay(n // graph_steps)
            b.record()
        else:
            for _ in range(5):
                env.prepare()
                env.step_random()
            torch.cuda.synchronize()
            a.record()
            for _ in range(n):
                env.prepare()
                env.step_random()
            b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    for name, kw, m, n, graph_steps in (
        ("core_only_replay_M2^24", dict(rng_mode="replay"), 1 << 24, 48, 0),
        ("core_only_philox_M2^24", dict(rng_mode="philox"), 1 << 24, 48, 0),
        ("fused_bf16_onehot_M2^24", dict(rng_mode="replay", onehot="bf16"), 1 << 24, 48, 0),
        ("fused_u8_onehot_M2^24", dict(rng_mode="replay", onehot="u8"), 1 << 24, 48, 0),
        ("train_shape_M2048_fused_f32_eager", dict(rng_mode="replay", onehot="f32"), 2048, 192, 0),
        ("train_shape_M2048_fused_f32_graph16", dict(rng_mode="replay", onehot="f32"), 2048, 192, 16),
        ("train_shape_M4096_fused_f32_graph16", dict(rng_mode="replay", onehot="f32"), 4096, 192, 16),
    ):
        env = ml2048_b200.VecGame(m, ml2048_b200.reward_fn_improved if m < 100000 else None, output="torch",
                                  track_merged=False, sync_free=True, device=dev, **kw)
        env.reset(seed)
        for _ in range(64 if m > 100000 else 128):
            env.prepare()
            env.step_random()
        ms = timed(env, n, graph_steps)
        out[name] = {"us_per_step": ms * 1e3, "env_steps_per_s": m / (ms * 1e-3)}
        if name == "core_only_replay_M2^24":
            # the environment alone: actions GIVEN (recorded from this trajectory, device-resident), step kernel timed by itself
            snap = env.state_dict()
            k = 24
            acts = torch.empty((k, m), dtype=torch.uint8, device=dev)
            for t in range(k):
                env.prepare()
                env.step_random(return_actions=True)
                acts[t].copy_(env.sampled_actions)
            env.load_state_dict(snap)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
            for t in range(k):
                env.prepare()
                ev[t][0].record()
                env.step(acts[t])
                ev[t][1].record()
            torch.cuda.synchronize()
            kms = sum(a.elapsed_time(b) for a, b in ev[4:]) / (k - 4)
            out["core_only_given_actions_step_kernel_M2^24"] = {
                "us_per_launch": kms * 1e3, "env_steps_per_s": m / (kms * 1e-3),
                "hbm_frac": BYTES_CORE * m / (kms * 1e-3) / 1e9 / load_peaks()[0]}
            del snap, acts
        del env
        torch.cuda.empty_cache()
    return out


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the e2e arm are
    allocated (first touch) on the NUMA node the GPU's PCIe root hangs off.  Harmless when it cannot be done."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to GPU {local_rank}"
    except Exception as exc:  # noqa: BLE001
        return f"not bound ({type(exc).__name__})"
    return "not bound"


def main() -> None:
    args = parse_args()
    # exactly ONE line may reach stdout (the driver parses it): everything libraries print to fd 1 (NCCL's version
    # banner, for one) is sent to stderr, and the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
