#!/usr/bin/env python
"""bench.py -- fixed-wing env-steps/sec on B200 (BASELINE.json metric), one JSON line on stdout.

Our arm (default):
  a "step" is ONE launch of the fused env-step kernel over one batch of `--envs` environments per GPU
  (default 65,536 = BASELINE configs[1], "batched fixed-wing physics step only, random actions"):
  8 physics substeps at 240 Hz per env, actions drawn in-kernel from Philox, motor noise on,
  auto-reset on ground/dome.  `value` = envs * K * n_gpus / device time (CUDA events on the launching
  stream, barrier + synchronize on both sides, MAX over ranks).
  L2 hygiene: the per-batch state (~7 MB) would sit in the 126 MB L2 between launches, so the timed loop
  rotates over enough independent env batches that the working set exceeds 2x L2 ("l2" in config).
  `e2e` = the same metric through the reference-facing host call: FixedwingVecEnv.step_arrays(actions)
  (the VecEnv.step seam; C ABI fw_step_host) on the Fixedwing-Waypoints-v3 task with HOST buffers:
  H2D of the actions and D2H of obs/reward/flags inside the timed region.
  `roofline`   : dominant kernel fw_step_kernel; algorithmic bytes (152 B/env-step physics-only)
                 over its mean launch duration against the measured HBM peak, plus the FP32-pipe view that
                 actually binds it (6,400 flop/env-step against a live-measured FMA-chain peak).
  `cpu_baseline`: the fp64 oracle (kind "port") on the host cores, bounded sample, rank 0 at N=1 only.

Reference arm (--impl reference): the reference's CPU path for this metric.  PyFlyt/pybullet cannot be
installed offline and the reference has no compilable sources, so this times the oracle port of the same
semantics on ALL host cores (bounded sample per step); rank 0 only under torchrun.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fixedwing_env_steps_per_sec"
UNIT = "env-steps/s"
# SURVEY.md section 8(d); lowlevel (2 substeps per env-step): state 76 in + 76 out, action 24, target 12, obs 84, reward 4
# objlock_duck: state 76 + 76, action 16, wind 32, duck/vision planes 80 + 80, vision history 124 + 124, obs 224, reward 4
BYTES_PER_STEP = {"physics_only": 152, "waypoints_v3": 332, "waypoint_objlock": 400, "lowlevel": 276, "objlock_duck": 836}
FLOPS_PER_STEP = {"physics_only": 6400, "waypoints_v3": 7000, "waypoint_objlock": 7000, "lowlevel": 1750, "objlock_duck": 7000}
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=65536, help="environments per GPU per launch")
    ap.add_argument("--workload", choices=["physics_only", "waypoints_v3", "waypoint_objlock", "lowlevel", "objlock_duck", "ppo"],
                    default="physics_only",
                    help="ppo = BASELINE configs[2]: PPO Fixedwing-Waypoints rollout+update (a step is one PPO iteration)")
    ap.add_argument("--ppo-preset", choices=["waypoints_v3", "waypoint_objlock", "lowlevel", "objlock_duck"], default="waypoints_v3")
    ap.add_argument("--ppo-envs", type=int, default=4096)
    ap.add_argument("--ppo-n-steps", type=int, default=128)
    ap.add_argument("--ppo-minibatches", type=int, default=4)
    ap.add_argument("--ppo-epochs", type=int, default=20)
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--steps-per-launch", type=int, default=1,
                    help="env-steps fused into one launch (state kept in registers); a bench step is one launch")
    ap.add_argument("--no-graph", action="store_true", help="issue launches from the host loop instead of a CUDA graph")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons during the timed region (the quantities of the B200_PROFILING.md nvidia-smi recipe,
    read through NVML when nvidia_ml_py is importable, else by polling nvidia-smi)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def _nvml(self):
        """Fast path: NVML through nvidia_ml_py (sub-millisecond per sample, so a 40 ms timed region still gets tens of
        samples); None when the module or the device is not available -> nvidia-smi polling."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                    ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                    ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))

            def sample():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                return [str(sm), str(mx), "0"] + ["Active" if (r & b) else "Not Active" for _, b in bits]
            sample()
            return sample
        except Exception:
            return None

    def run(self):
        sample = self._nvml()
        while not self._stop_evt.is_set():
            if sample is not None:
                try:
                    self.samples.append(sample())
                except Exception:
                    pass
                self._stop_evt.wait(0.002)
                continue
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            # a timed region shorter than one polling interval: read the clocks once, right after it
            sample = self._nvml()
            try:
                if sample is not None:
                    self.samples.append(sample())
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for p in self.samples:
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def workload_name(workload: str, n: int) -> str:
    """config.workload, identical for our arm and the reference arm."""
    if workload == "physics_only":
        return (f"{workload}: BASELINE configs[1] batched fixed-wing physics step, {n} envs/GPU, in-kernel Philox random "
                f"actions, motor noise on")
    return f"{workload}: {n} envs/GPU, random actions"


def oracle_rate(workload: str, threads: int, seconds: float, n_envs: int = 4096):
    """Time the fp64 oracle (oracle/) on a bounded sample of the same workload: random actions from the
    same Philox stream, same auto-reset.  Returns (env-steps/s, sample description)."""
    from oracle import fw_oracle as fo
    import pyflyt_drone_b200 as fw
    cfg = fw.make_config(workload)
    env = fo.OracleVecEnv(cfg.as_dict(), n_envs, seed=0, nthreads=threads)
    env.reset()
    t0 = time.perf_counter()
    env.rollout_random(2)
    probe = (time.perf_counter() - t0) / 2
    steps = max(2, min(20000, int(seconds / max(probe, 1e-6))))
    t0 = time.perf_counter()
    done = env.rollout_random(steps)
    dt = time.perf_counter() - t0
    return done / dt, f"{n_envs} envs x {steps} agent steps ({done} env-steps, {dt:.1f} s, {threads} threads, fp64 oracle port)"


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = host_cores()
    from oracle import fw_oracle as fo
    import pyflyt_drone_b200 as fw
    cfg = fw.make_config(args.workload)
    n_envs = 4096
    env = fo.OracleVecEnv(cfg.as_dict(), n_envs, seed=0, nthreads=threads)
    env.reset()
    # bound the run: K "steps" of the reference arm are K agent steps of the sample batch, capped to ~60 s
    t0 = time.perf_counter(); env.rollout_random(1); per = time.perf_counter() - t0
    K = max(1, min(args.steps, int(60.0 / max(per, 1e-6))))
    W = max(1, min(args.warmup, max(1, int(5.0 / max(per, 1e-6)))))
    env.rollout_random(W)
    t0 = time.perf_counter()
    done = env.rollout_random(K)
    dt = time.perf_counter() - t0
    v = done / dt
    sample = f"{n_envs} envs per step, {K} timed steps ({done} env-steps), {threads} host threads, {cpu_model()}"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.envs), "envs_per_gpu": args.envs,
                   "sample_envs_per_step": n_envs,
                   "note": "reference CPU path restated as the fp64 oracle port (PyFlyt/pybullet are not installable "
                           "offline); each step is a bounded sample of the workload: one agent step of a "
                           f"{n_envs}-env slice of the batch on all host threads"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the env step has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import ctypes as C
    import pyflyt_drone_b200 as fw
    from pyflyt_drone_b200 import _lib
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv

    over = {}
    cfg = fw.make_config(args.workload, **over)
    N = args.envs
    state_bytes = N * (6 * 16 + 4 + (cfg.num_targets * 12) + (5 * 16 + 32 * 12 if cfg.task in (2, 4) else 0)
                       + (31 * 4 if cfg.task == 4 else 0))
    replicas = max(2, int(np.ceil(2 * L2_BYTES / state_bytes)))
    envs = [FixedwingVecEnv(N, config=cfg, device=local_rank, seed=1234, env_id0=(rank * replicas + r) * N)
            for r in range(replicas)]
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spl = max(1, args.steps_per_launch)
    FixedwingVecEnv.rollout_random(envs, W, spl, use_graph=False)
    if not args.no_graph:
        # one untimed pass through the launch graph: its construction / instantiation is not part of a step
        FixedwingVecEnv.rollout_random(envs, replicas, spl, use_graph=True)
        if K % replicas:
            FixedwingVecEnv.rollout_random(envs, K % replicas, spl, use_graph=True)     # and the graph of the tail
    launches0 = sum(e.launch_count for e in envs)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    FixedwingVecEnv.rollout_random(envs, K, spl, use_graph=not args.no_graph)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None       # sampled during the timed region only
    ms = e0.elapsed_time(e1)
    launches = sum(e.launch_count for e in envs) - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = N * K * spl * world / (ms * 1e-3)
    kernel_ms = ms / K          # back-to-back launches of one kernel: the mean launch-to-launch duration

    # ---- e2e: the VecEnv.step seam with host buffers (Waypoints-v3 task, obs/reward/flags come back) ----
    e2e_cfg = fw.make_config("waypoints_v3", **over)
    venv = FixedwingVecEnv(N, config=e2e_cfg, device=local_rank, seed=99, env_id0=rank * N)
    venv.reset()
    rng = np.random.default_rng(rank)
    # the step's inputs wait in pinned host memory (bench contract); FixedwingVecEnv reads page-locked caller buffers in place
    pinned = [torch.empty((N, 4), dtype=torch.float32, pin_memory=True) for _ in range(4)]   # owns the memory
    acts = [b.numpy() for b in pinned]
    for a in acts:
        a[:] = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    for s in range(5):
        venv.step_arrays(acts[s % 4], want_terminal_obs=False)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.e2e_steps):
        venv.step_arrays(acts[s % 4], want_terminal_obs=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N * args.e2e_steps * world / float(t.item())
    h2d = N * 4 * 4
    d2h = N * venv.obs_dim * 4 + N * 4 + N

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        tf, sms, khz = C.c_double(), C.c_int32(), C.c_int32()
        _lib.check(_lib.load().fw_measure_fp32_peak(local_rank, C.byref(tf), C.byref(sms), C.byref(khz)))
        per_gpu = N * spl / (kernel_ms * 1e-3)
        achieved_gbs = per_gpu * BYTES_PER_STEP[args.workload] / 1e9
        fp32_tf = per_gpu * FLOPS_PER_STEP[args.workload] / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"{args.workload}_{N}")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, N),
                       "envs_per_gpu": N, "substeps_per_env_step": cfg.inner_per_step * cfg.substeps_per_inner,
                       "physics_substeps_per_sec": value * cfg.inner_per_step * cfg.substeps_per_inner,
                       "l2": f"rotating {replicas} env batches ({replicas * state_bytes / 2**20:.0f} MiB > 2x L2)",
                       "env_steps_per_launch": spl,
                       "launch": "host loop" if args.no_graph else f"CUDA graph of {replicas} launches", "e2e_workload": "waypoints_v3 via FixedwingVecEnv.step_arrays (actions in pinned host numpy arrays, obs/reward/flags back in host memory)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "fw_step_kernel", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_env_step": BYTES_PER_STEP[args.workload],
                         "binding_pipe": "fp32",
                         "fp32": {"achieved": fp32_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": fp32_tf / tf.value,
                                  "flops_per_env_step": FLOPS_PER_STEP[args.workload], "sm_count": sms.value,
                                  "peak_source": "FMA chain measured in this run (fw_measure_fp32_peak)"}},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, sample = oracle_rate(args.workload, host_cores(), args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": host_cores(), "kind": "port",
                                    "sample": sample + f"; {cpu_model()}"}
        print(json.dumps(line), flush=True)
    for e in envs:
        e.close()
    venv.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ppo(args, rank: int, local_rank: int, world: int):
    """BASELINE configs[2]: PPO on Fixedwing-Waypoints (train_Fixedwing_Waypoints_v3 hyper-parameters, scaled
    variant B of SURVEY 8d: n_steps 128, 4 minibatches per epoch).  A step = one rollout + one update."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pyflyt_drone_b200.ppo import PPO
    from pyflyt_drone_b200.vec_env import FixedwingVecEnv
    N, T = args.ppo_envs, args.ppo_n_steps
    env = FixedwingVecEnv(N, preset=args.ppo_preset, device=local_rank, seed=42, env_id0=rank * N)
    model = PPO("MlpPolicy", env, learning_rate=3e-4, n_steps=T, batch_size=N * T // args.ppo_minibatches,
                n_epochs=args.ppo_epochs, gamma=0.99, gae_lambda=0.95, clip_range=0.2, ent_coef=0.001, vf_coef=0.5,
                max_grad_norm=0.5, seed=42)
    K, W = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    model.learn(W * N * T * world)
    model.stats.__init__()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    model.learn(K * N * T * world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        st = model.stats
        line = {"metric": "ppo_env_steps_per_sec", "value": K * N * T * world / dt, "unit": "env-steps/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"ppo: {args.ppo_preset} PPO rollout+update (BASELINE configs[2]/[3])",
                           "envs_per_gpu": N, "n_steps": T, "minibatches_per_epoch": args.ppo_minibatches,
                           "n_epochs": args.ppo_epochs, "rollout_s": st.rollout_s, "update_s": st.update_s,
                           "rollout_env_steps_per_sec": st.env_steps / max(st.rollout_s, 1e-9),
                           "update": "fused tcgen05 minibatch gradient kernel + clip/Adam kernel (csrc/ppo_update_tc.cu)"
                                     if model.update == "kernel" else "torch autograd (6-channel or > 32-float policies)",
                           "forward": "tcgen05 kind::tf32 policy/value forward (csrc/ppo_tc.cu)" if model.tensor_core_forward
                                      else "CUDA-core fp32 policy/value forward (csrc/ppo_kernels.cu)", "rollout": "CUDA graph"},
                # the rollout is one CUDA-graph replay, which bypasses the C-side launch counter: count this repo's kernels
                # from the launch sequence instead -- per rollout step: obs moments, policy forward, env step, return
                # moments, reward finalise, bootstrap value forward, step counter; per rollout: last values, GAE; per
                # epoch: permutation; per minibatch: advantage stats, gradient, partial reduce, clip+Adam
                "gpu_launches": int(K * (T * 7 + 2 + args.ppo_epochs * (1 + args.ppo_minibatches * 4)))}
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "ppo":
        run_ppo(args, rank, local_rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
