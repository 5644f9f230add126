#!/usr/bin/env python
"""bench.py — env-steps/s of the tinycarlo hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

Workload (config.workload): BASELINE.json configs[2] — Knuffingen map, 480x640 `classes` observations (5x480x640 u8 =
1 536 000 B per env-step), Stanley-controller actions with lanepath CTE / heading info consumed every step, auto-reset
of finished envs; 16384 envs per GPU (weak scaling: per-GPU work is fixed, envs shard by index, no per-step collective;
NCCL only all-gathers episode statistics).

  python bench.py --gpus 1 --steps 100 --warmup 5                      # this repo's CUDA path
  python bench.py --impl reference --steps 3 --warmup 1                # the CPU arm: oracle port on all host cores
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one lockstep pass of the hot path over all envs of the rank. `value` is measured with the action tensors
already on the device (CUDA events on the launching stream, max over ranks); `e2e` drives the same step through the
host-buffer entry point (pinned host actions in, scalar results out, observations stay device-resident as the
vectorised API defines); `roofline` is the rasterise+store kernel's algorithmic bytes over its own event-timed duration.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

OBS_H, OBS_W, N_CLASSES = 480, 640, 5
OBS_BYTES = N_CLASSES * OBS_H * OBS_W
METRIC = "env-steps/sec (480x640 class obs, Knuffingen)"
UNIT = "env-steps/s"


def bench_config():
    from pair_util import make_config
    return make_config("knuffingen", "classes", cam={"resolution": [OBS_H, OBS_W]})


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                                          str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_arm(n_envs, steps, warmup, threads, min_seconds=0.0):
    """The reference's algorithm restated in C (oracle/tc_oracle.c), all host threads via OpenMP, same workload:
    Knuffingen 480x640 classes, Stanley actions from the info of the previous step. Runs `steps` lockstep passes, and keeps
    going until min_seconds have passed. Returns (env-steps/s, seconds, passes)."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import oracle as orc
    from pair_util import oracle_env, stanley_actions
    cfg = bench_config()
    oenv = oracle_env(cfg, n_envs)
    rng = [orc.make_rng(i) for i in range(n_envs)]
    oenv.reset([oenv.map.sample_spawn_node(r) for r in rng])
    man = np.zeros(n_envs, np.int32)

    def one():
        cc = stanley_actions(oenv.cte.copy(), oenv.heading_error.copy(), cfg["car"]["max_steering_angle"])
        oenv.step(cc.astype(np.float64), man)
        done = (oenv.terminated | oenv.truncated).astype(bool)
        if done.any():
            oenv.reset([oenv.map.sample_spawn_node(r) if d else 0 for r, d in zip(rng, done)], mask=done)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done_steps = 0
    while done_steps < steps or time.perf_counter() - t0 < min_seconds:
        one()
        done_steps += 1
    dt = time.perf_counter() - t0
    return n_envs * done_steps / dt, dt, done_steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=16384)
    ap.add_argument("--cpu-envs", type=int, default=0, help="envs of the CPU sample (default: 128 x host threads, at most 2048: 3 GB of frames)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="length of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fill-context", action="store_true", help="skip torch's zero-fill of the observation tensor (context number; keeps ncu launch lists clean)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workload = f"knuffingen 480x640 classes, Stanley actions + lanepath info, auto-reset, {args.envs_per_gpu} envs/GPU"

    if args.impl == "reference":
        # The reference is single-process pure Python and cannot travel to the GPU box; its algorithm restated in C
        # (the pinned oracle) runs on all host threads instead. Rank 0 only.
        if rank != 0:
            return
        # one "step" of this arm = one lockstep pass over a bounded sample of the workload (n envs instead of 16384 per GPU)
        n = args.cpu_envs or min(128 * threads, 2048)
        W = max(args.warmup, 1)
        val, dt, _ = cpu_arm(n, args.steps, W, threads)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": W,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": workload, "cpu_sample": f"{n} envs x {args.steps} steps"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{n} envs x {args.steps} lockstep steps ({dt:.1f} s), oracle/tc_oracle.c with OpenMP"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from tinycarlo_b200 import TinyCarloVecEnv
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = args.envs_per_gpu
    cfg = bench_config()
    env = TinyCarloVecEnv(cfg, N, device=dev, env_index_offset=rank * N, autoreset="next_step")
    max_steer = float(cfg["car"]["max_steering_angle"])
    speed, k_gain = 0.8, 4.0
    maneuver = torch.zeros(N, dtype=torch.int32, device=dev)
    cc = torch.zeros((N, 2), dtype=torch.float32, device=dev)
    cc[:, 0] = speed
    stats = torch.zeros(4, dtype=torch.float64, device=dev)  # episodes finished, truncations, reward sum, env-steps
    gathered = torch.zeros(world * 4, dtype=torch.float64, device=dev) if world > 1 else None

    env.reset(seed=0)

    def step_device():
        # examples/stanley_control.py:56-58 as tensor ops on the info of the previous step
        o = env.out
        cc[:, 1] = (o["heading_error"] + torch.atan2(k_gain * o["cte"], torch.full_like(o["cte"], speed))) * (180.0 / np.pi / max_steer)
        _, reward, term, trunc, _ = env.step({"car_control": cc, "maneuver": maneuver})
        stats[0] += term.sum()
        stats[1] += trunc.sum()
        stats[2] += reward.sum()
        stats[3] += N

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()   # sampled from the warm-up to the end of the device-timed region (all under load)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = env.launch_count
    env.profile_begin(args.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    if world > 1:
        dist.all_gather_into_tensor(gathered, stats)  # episode statistics over NVLink: the only collective
    e1.record()
    barrier()
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    kern_ms, kern_steps = env.profile_end()
    launches = env.launch_count - launches0
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = float(elapsed_ms.item()) / args.steps
    value = N * world / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer entry point (pinned host actions in, scalar results out)
    e2e = None
    if not args.no_e2e:
        h_cc = torch.zeros((N, 2), dtype=torch.float32).pin_memory()
        h_man = torch.zeros(N, dtype=torch.int32).pin_memory()
        h_rew = torch.zeros(N, dtype=torch.float32).pin_memory()
        h_term = torch.zeros(N, dtype=torch.uint8).pin_memory()
        h_trunc = torch.zeros(N, dtype=torch.uint8).pin_memory()
        h_cte = torch.zeros(N, dtype=torch.float32).pin_memory()
        h_head = torch.zeros(N, dtype=torch.float32).pin_memory()
        h_cc[:, 0] = speed
        cc_np, cte_np, head_np = h_cc.numpy(), h_cte.numpy(), h_head.numpy()

        def step_host():
            cc_np[:, 1] = (head_np + np.arctan2(k_gain * cte_np, speed)) * (180.0 / np.pi / max_steer)
            env.step_host(h_cc, h_man, h_rew, h_term, h_trunc, h_cte, h_head)
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        per_step = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            step_host()
            per_step.append(time.perf_counter() - ts)
        torch.cuda.synchronize(dev)
        if os.environ.get("TC_BENCH_DEBUG") and rank == 0:
            print("e2e per-step ms:", [round(x * 1e3, 2) for x in per_step], file=sys.stderr)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": N * world * args.steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": N * 12,
               "d2h_bytes_per_step": N * 14, "note": "tc_step_host: pinned host actions in, reward/terminated/truncated/cte/heading out; "
               "observations stay in HBM (the vectorised entry point returns CUDA tensors)"}

    # for context next to the copy peak: torch's own fill kernel zeroing the same observation tensor
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    write_ceiling = 0.0
    for _ in range(0 if args.no_fill_context else 3):
        w0.record()
        env.obs.zero_()
        w1.record()
        torch.cuda.synchronize(dev)
        write_ceiling = max(write_ceiling, env.obs.numel() / (w0.elapsed_time(w1) * 1e-3) / 1e9)
    # ... and the driver's own memset of the same tensor (cudaMemsetAsync through libcudart): the pure-write ceiling of this box
    memset_gbs = None
    if not args.no_fill_context:
        try:
            import ctypes
            rt = None
            for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
                try:
                    rt = ctypes.CDLL(name)
                    break
                except OSError:
                    continue
            if rt is not None:
                rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
                nbytes = env.obs.numel() * env.obs.element_size()
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                best = 0.0
                for _ in range(3):
                    w0.record()
                    rc = rt.cudaMemsetAsync(ctypes.c_void_p(env.obs.data_ptr()), 0, nbytes, stream)
                    w1.record()
                    torch.cuda.synchronize(dev)
                    if rc == 0:
                        best = max(best, nbytes / (w0.elapsed_time(w1) * 1e-3) / 1e9)
                memset_gbs = best or None
        except Exception:
            memset_gbs = None
    # ---- the same host loop WITH the observations copied to pinned host memory every step (what the single-env gymnasium
    # drop-in does), on a reduced batch so that it stays a measurement of the path and not of minutes of PCIe: rank 0 only
    e2e_obs = None
    if not args.no_e2e and rank == 0:
        n_small = min(N, 512)
        env_s = TinyCarloVecEnv(cfg, n_small, device=dev, autoreset="next_step")
        env_s.reset(seed=0)
        h_obs = torch.zeros(env_s.obs.shape, dtype=torch.uint8).pin_memory()
        hs = [torch.zeros((n_small, 2), dtype=torch.float32).pin_memory(), torch.zeros(n_small, dtype=torch.int32).pin_memory(),
              torch.zeros(n_small, dtype=torch.float32).pin_memory(), torch.zeros(n_small, dtype=torch.uint8).pin_memory(),
              torch.zeros(n_small, dtype=torch.uint8).pin_memory(), torch.zeros(n_small, dtype=torch.float32).pin_memory(),
              torch.zeros(n_small, dtype=torch.float32).pin_memory()]
        hs[0][:, 0] = speed

        def step_host_obs():
            hs[0].numpy()[:, 1] = (hs[6].numpy() + np.arctan2(k_gain * hs[5].numpy(), speed)) * (180.0 / np.pi / max_steer)
            env_s.step_host(*hs)
            h_obs.copy_(env_s.obs, non_blocking=True)
            torch.cuda.synchronize(dev)
        for _ in range(2):
            step_host_obs()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host_obs()
        dt_s = time.perf_counter() - t0
        e2e_obs = {"value": n_small * args.steps / dt_s, "unit": UNIT, "envs": n_small, "h2d_bytes_per_step": n_small * 12,
                   "d2h_bytes_per_step": n_small * 14 + int(h_obs.numel()),
                   "note": "as e2e, plus the u8 observations copied to pinned host memory every step (PCIe-bound)"}
        env_s.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rasterise + store): algorithmic bytes = the observation, written once
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    raster_ms = kern_ms["raster"] / max(kern_steps, 1)
    achieved = N * OBS_BYTES / (raster_ms * 1e-3) / 1e9 if raster_ms > 0 else 0.0
    # DRAM traffic of that kernel from the committed ncu capture (profiles/r01_traffic.json: dram__bytes_read.sum +
    # dram__bytes_write.sum of one launch, per env), scaled to this launch's env count; null if the file is missing
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f)["render_kernel_dram_bytes_per_env"] * N
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes": N * OBS_BYTES,
                "kernel": "tc_render_classes_kernel", "kernel_ms_per_launch": raster_ms,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "step_share_ms": {k: v / max(kern_steps, 1) for k, v in kern_ms.items()},
                "frac_of_nominal_8tbs": achieved / 8000.0,
                "torch_zero_fill_gbs": write_ceiling if write_ceiling > 0 else None, "cuda_memset_gbs": memset_gbs}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        n_cpu = args.cpu_envs or min(128 * threads, 2048)
        val, dt, passes = cpu_arm(n_cpu, 8, 1, threads, min_seconds=args.cpu_seconds)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n_cpu} envs x {passes} lockstep steps of the same workload, oracle/tc_oracle.c with OpenMP ({dt:.1f} s)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "obs_bytes_per_env_step": OBS_BYTES, "envs_total": N * world,
                       "l2": "observation tensor (25 GB/GPU) is far larger than L2; nothing is re-read between steps",
                       "obs_gbs": value * OBS_BYTES / 1e9},
            "clocks": clocks, "e2e": e2e, "e2e_obs_to_host": e2e_obs, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "episode_stats": {"finished": float(stats[0].item()), "truncated": float(stats[1].item()), "reward_sum": float(stats[2].item())}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
