#!/usr/bin/env python
"""bench.py -- humanoid physics env-steps/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], largest single-GPU size): humanoid_mjx.xml, 262144 envs per GPU, trajectory
distribution B of SURVEY.md 8d (reset noise per src/envs.py:127-131,147, i.i.d. clip(N(0,1)) actions, auto-reset).
A "step" is one pass of the hot path over the batch: v_step fused with the trainer's auto-reset (train_ppo.py:143-161).
  value : env-steps/s summed over all ranks, inputs resident in HBM (actions / reset keys pre-generated on the device)
  e2e   : the same step through the C ABI's host-buffer entry point (mjxb_step_autoreset_host): pinned host actions+keys
          -> H2D -> kernel -> D2H of obs/reward/terminated/truncated, inside the timed region
  roofline     : HBM roofline of the step kernel on its algorithmic bytes (1056 B / env-step)
  roofline_fp32: the bound that actually applies (FP32 CUDA cores): counted algorithmic FLOPs / measured kernel time
  cpu_baseline : the CPU oracle (restatement of the reference's MJX path; the reference itself cannot run here) on a
                 bounded sample of the same workload, all host threads
`--impl reference` times that CPU restatement as the reference arm (MJX / MuJoCo are not installable in this image).
"""
from __future__ import annotations

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

METRIC = "humanoid_env_steps_per_sec"
UNIT = "env-steps/s"
ALG_BYTES_PER_STEP = 1056          # state r/w 2*368 + action 84 + obs/reward/done 228 + key 8
FP32_NOMINAL_TFLOPS = 74.4         # 148 SM x 128 lanes x 2 x 1.965 GHz (nominal; the line reports the MEASURED FFMA peak beside it)


def counted_flops(oracle, qpos, qvel, warm, ctrl):
    """Algorithmic FLOPs per env-step COUNTED by the CPU oracle's stage counters on a sample of the measured state distribution
    (oracle/oracle.hpp: `flops` = the dense formulation MJX executes, `flops_act` = the activity-aware minimum of SURVEY.md 8d:
    candidate rows only, rows active at each Newton iterate in J^T D J, tree-sparse factor_m; counting rules in DESIGN.md)."""
    out = oracle.physics_step(qpos, qvel, warm, None, ctrl, prec="f32", integrate=False,
                              debug=("flops", "flops_act", "solver_niter", "efc_active"))
    act = (out["efc_active"] >> 1).sum(1)
    cand = (out["efc_active"] & 1).sum(1)
    return dict(flops_per_env_step=float(out["flops_act"].mean()), flops_per_env_step_dense=float(out["flops"].mean()),
                mean_newton_iters=float(out["solver_niter"].mean()), mean_candidate_rows=float(cand.mean()),
                mean_active_rows=float(act.mean()), sample_envs=int(qpos.shape[0]))


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max(float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()) if self.rows else None,
                "samples": len(self.rows), "reasons": sorted(reasons)}


def cpu_reference_run(n_env: int, steps: int, warmup: int, seed: int = 42):
    """The CPU restatement of the reference path (oracle/), all host threads, on a bounded sample of the bench workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from oracle import oracle as O
    O.build()
    model = helpers.load()
    orc = helpers.make_oracle(model, helpers.env_config())
    # every host core this process may use, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    orc.nthreads = ncpu
    rng = np.random.default_rng(seed)
    st, _ = orc.env_reset(helpers.ppo_keys(seed, n_env))
    for t in range(warmup):
        st, *_ = orc.env_step(st, np.clip(rng.normal(size=(n_env, 21)), -1, 1), reset_keys=helpers.ppo_keys(1000 + t, n_env))
    t0 = time.perf_counter()
    for t in range(steps):
        st, *_ = orc.env_step(st, np.clip(rng.normal(size=(n_env, 21)), -1, 1), reset_keys=helpers.ppo_keys(2000 + t, n_env))
    dt = time.perf_counter() - t0
    return n_env * steps / dt, dt, ncpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs", type=int, default=262144, help="envs per GPU (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-envs", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--ppo-large", action="store_true", help="(default now) also time config 5 (65536 envs/GPU PPO iteration)")
    ap.add_argument("--no-ppo-large", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-apg", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"humanoid_mjx v_step+auto-reset, {args.envs} envs/GPU, trajectory distribution B (reset noise, clip(N(0,1)) actions)",
              "model_xml": "models/humanoid_mjx.xml", "envs_per_gpu": args.envs, "solver": "Newton 10/20 pyramidal, implicitfast, dt=0.005",
              "parallelism": f"env-sharded x{world}, no data-path collective",
              "l2": "state+io per step = 0.28 GB > 126 MB L2 (inputs larger than L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        n = min(args.cpu_sample_envs, args.envs)
        steps = max(1, min(args.steps, 8))
        val, dt, threads = cpu_reference_run(n, steps, min(args.warmup, 3))
        config = dict(config, envs_per_gpu=n, steps_timed=steps,
                      workload=f"humanoid_mjx v_step+auto-reset, BOUNDED SAMPLE of the bench workload: {n} envs x {steps} steps on {threads} host "
                               f"threads (the GPU arm runs {args.envs} envs/GPU), trajectory distribution B",
                      l2="n/a (CPU)", parallelism=f"OpenMP over envs, {threads} threads")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3),
                "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                 "sample": f"{n} envs x {steps} steps of the bench workload (CPU restatement of mjx.step + src/envs.py; "
                                           "mujoco-mjx / jax are not installable in this image)"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    from mujoco_mjx_lab_b200 import _lib, mjx, modelc, parallel, training_utils
    from mujoco_mjx_lab_b200.config import EnvConfig

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    numa_cpus = parallel.bind_to_gpu_numa(local_rank) if world > 1 and os.environ.get("MJXB_NO_NUMA_BIND") is None else 0
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        parallel.init("nccl")
    model = modelc.builtin_model("humanoid_mjx")
    cfg = EnvConfig(posture_penalty_weight=0.0, random_flip=True)        # effective PPO defaults (src/config.json)
    m, sysm, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step = training_utils.load_model_and_create_env("", cfg, model=model)
    env_sys = v_step.sys
    n = args.envs
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    nbuf = 8
    acts = [torch.randn(n, nu, device=dev, generator=g).clamp_(-1, 1) for _ in range(nbuf)]
    keys = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device=dev, dtype=torch.int32, generator=g) for _ in range(nbuf)]
    state, obs = v_reset(torch.from_numpy(parallel.rank_keys(42, rank, n).view(np.int32)).to(dev))
    # settle into the trajectory distribution (envs fall and get reset at different times) before measuring
    for i in range(60):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % nbuf], keys[i % nbuf], inplace=True)
    for i in range(args.warmup):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % nbuf], keys[i % nbuf], inplace=True)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    parallel.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L = _lib.lib()
    launches_0, launch_calls = L.mjxb_launch_count(), args.steps
    e0.record()
    for i in range(args.steps):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % nbuf], keys[i % nbuf], inplace=True)
    e1.record()
    gpu_launches = int(L.mjxb_launch_count() - launches_0)
    torch.cuda.synchronize()
    parallel.barrier()
    ms_total = parallel.max_over_ranks(e0.elapsed_time(e1), dev)
    if sampler:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    ms_step = ms_total / args.steps
    value = n * world * args.steps / (ms_total * 1e-3)
    done_frac = float(torch.maximum(te, tr).mean())

    # ---- end to end through the host-buffer ABI call (pinned host buffers; H2D + kernel + D2H timed)
    h = env_sys.handle
    pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, pin_memory=True)
    h_act = [pin(n, nu) for _ in range(nbuf)]                        # the same action / key sequence as the device-resident leg
    h_keys = [pin(n, 2, dt=torch.int32) for _ in range(nbuf)]
    for b, src in zip(h_act + h_keys, acts + keys):
        b.copy_(src.cpu())
    h_obs, h_r, h_te, h_tr = pin(n, env_sys.obs_dim), pin(n), pin(n), pin(n)
    d, aux = state
    host_state = [np.ascontiguousarray(t.detach().cpu().numpy()) for t in (d.qpos, d.qvel, d.qacc_warmstart, d.time, aux)]  # keep alive
    _lib.check(L.mjxb_state_set_host(h, n, *[a.ctypes.data for a in host_state]), "state_set_host")
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step(i):
        _lib.check(L.mjxb_step_autoreset_host(h, n, h_act[i % nbuf].data_ptr(), h_keys[i % nbuf].data_ptr(), h_obs.data_ptr(), h_r.data_ptr(),
                                              h_te.data_ptr(), h_tr.data_ptr()), "step_autoreset_host")
    for i in range(3):
        e2e_step(i)
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = parallel.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_value = n * world * e2e_steps / e2e_s
    h2d, d2h = n * (nu * 4 + 8), n * (env_sys.obs_dim + 3) * 4

    # ---- solver statistics of the measured distribution (GPU, untimed) and a sample of it for the oracle's FLOP counters
    nstat = min(n, 65536)
    dd = mjx.Data(d.qpos[:nstat], d.qvel[:nstat], d.qacc_warmstart[:nstat], d.time[:nstat], acts[0][:nstat])
    _, dbg = mjx.forward(env_sys, dd, debug=True)
    mean_iters = float(dbg["solver_niter"].float().mean())
    mean_rows = float((dbg["efc_active"] & 1).sum(1).float().mean())
    mean_active = float((dbg["efc_active"] >> 1).sum(1).float().mean())
    spill_frac = float(((dbg["status"] & 2) != 0).float().mean())
    nflop = min(n, 4096)
    flop_sample = [t[:nflop].double().cpu().numpy() for t in (d.qpos, d.qvel, d.qacc_warmstart, acts[0])]

    # ---- FP32 FMA-pipe ceiling of this GPU, measured (FFMA-saturating microbenchmark inside libmjxb.so)
    import ctypes as C
    ffma_tf, ffma_ms = C.c_float(0), C.c_float(0)
    _lib.check(L.mjxb_ffma_peak(local_rank, C.byref(ffma_tf), C.byref(ffma_ms)), "mjxb_ffma_peak")

    # ---- BASELINE configs[0..1]: the batch sizes the reference itself runs (64 = its CPU case, 4096 = mjx_humanoid_speed_test.py:141,
    # 1024 / 2048 = its PPO batches) and the rest of the 1K-256K sweep of configs[1] (the headline above is its 262,144-env point), both
    # input distributions of SURVEY.md 8d, on this rank's GPU
    sweep = []
    if not args.no_sweep:
        del state, obs
        for ns in (64, 1024, 2048, 4096, 16384, 65536):
            sweep.append(sweep_point(ns, v_reset, v_step, env_sys, nu, dev))

    # ---- PPO iteration time (the second half of BASELINE.json's metric): the caller of the hot path, reference train_ppo.py
    ppo = {}
    if not args.no_ppo:
        from mujoco_mjx_lab_b200 import ppo as ppo_mod
        del acts, keys
        torch.cuda.empty_cache()
        n3 = max(1024 // world, 16)                                  # config 3: 1024 envs total x 256 steps, sharded
        ppo["config3"] = ppo_mod.time_ppo(n3, 256, iters=5, warmup=3, minibatch_size=65536, model=model)
        if not args.no_ppo_large:                                    # config 5: 65536 envs per GPU, NCCL gradient all-reduce
            ppo["config5"] = ppo_mod.time_ppo(65536, 256, iters=2, warmup=1, minibatch_size=65536, model=model)
    apg = {}
    if not args.no_apg:                                              # config 4: APG 2048 envs x 128 horizon, CG 4/4, reverse-mode step kernels
        try:
            from mujoco_mjx_lab_b200 import apg as apg_mod
            apg["config4"] = apg_mod.time_apg(max(2048 // world, 16), 128, iters=3, warmup=2)
        except ImportError:
            apg["config4"] = {"unavailable": "apg module not built"}
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None                                                # dram bytes per step launch from the committed ncu capture
    try:
        tp = [q for q in (os.path.join(ROOT, "profiles", f) for f in ("r2d_traffic.json", "r2_traffic.json", "r1_traffic.json")) if os.path.exists(q)]
        tj = json.load(open(tp[0]))
        if int(tj["n_env"]) == n:
            traffic = float(tj["dram_bytes_per_launch"])
    except Exception:
        pass
    kernel_ms = ms_step                                           # one step == one step-kernel launch (+ two empty overflow passes)
    ach_gbs = ALG_BYTES_PER_STEP * n / (kernel_ms * 1e-3) / 1e9
    ffma_peak = float(ffma_tf.value)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "numa_bound_cpus": numa_cpus,
                "api": "mjxb_step_autoreset_host (pinned host buffers; one launch, action/keys streamed in behind ready flags, outputs stored into the caller's mapped buffers)"},
        "gpu_launches": gpu_launches,
        "gpu_launches_note": "counted by libmjxb.so (mjxb_launch_count) over the timed region: per step the schedule sort (batches >= 16,384 envs), the main tier and two overflow tiers that exit at once when their list is empty",
        "kernels": ["mjxb_sort_work_kernel (work-sorted schedule: envs ordered by descending cost of their previous step, 2048-env segments)",
                    "mjxb_step_kernel<false,32,16,16,true,true,true> (step: 32-row tile, 16 env-warps per SM, single-step instantiation, env groups taken from a device-wide counter)",
                    "mjxb_step_kernel<false,64,24,10,true,false> and <false,320,176,3,true,false> (overflow tiers; exit at once when their list is empty)"],
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": ALG_BYTES_PER_STEP * n, "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "algorithmic_bytes_per_env_step": ALG_BYTES_PER_STEP, "kernel_ms": kernel_ms,
                     "note": "the step is FP32-pipe/latency bound, ~16x under its HBM ceiling; see roofline_fp32"},
        "workload_stats": {"done_frac_per_step": done_frac, "overflow_rerun_frac": spill_frac, "mean_newton_iters_gpu": mean_iters,
                           "mean_candidate_rows_gpu": mean_rows, "mean_active_rows_gpu": mean_active},
        "published_reference_steps_per_sec": 72618, "x_published_reference": value / world / 72618.0,
        "clocks": sampler.summary() if sampler else None,
        "sweep": sweep,
        "ppo_iter": ppo,
        "apg_iter": apg,
        "arithmetic": "product build: --use_fast_math + closed-form exact line search where MJX iterates its own to convergence; both deviations "
                      "from MJX's float32 arithmetic are quantified against the reference-arithmetic build in profiles/r2_parity.json",
    }
    if not args.no_cpu_baseline and world == 1:
        ncpu = min(args.cpu_sample_envs, n)
        val, dt, threads = cpu_reference_run(ncpu, 4, 1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{ncpu} envs x 4 steps of the same workload, CPU oracle (float32 restatement of mjx.step + src/envs.py), {dt:.1f} s"}
    # BASELINE configs[0]: the reference's own CPU-runnable case -- mjx_humanoid_speed_test.py semantics at 64 envs -- on this box's host
    # cores with the CPU restatement (MJX-on-CPU itself is UNAVAILABLE here, see north_star_cpu_baselines), beside the GPU's N=64 sweep row
    if world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import helpers
            from oracle import oracle as O
            O.build()
            orc1 = helpers.make_oracle(model)
            try:
                orc1.nthreads = len(os.sched_getaffinity(0))
            except AttributeError:
                orc1.nthreads = os.cpu_count() or 1
            vel64 = np.linspace(0, 1, 64)
            orc1.speed_test(vel64, iters=2)
            t0 = time.perf_counter()
            it1 = 20
            orc1.speed_test(vel64, iters=it1)
            dt1 = time.perf_counter() - t0
            gpu64 = next((r for r in sweep if r["n_env"] == 64), None)
            line["config1_cpu_sanity"] = {"workload": "mjx_humanoid_speed_test.py:48-57 step, humanoid_mjx.xml, 64 envs", "cpu_env_steps_per_sec": 64 * it1 / dt1,
                                          "cpu_kind": "port (CPU restatement of mjx.step, float32)", "cores": orc1.nthreads, "iters": it1,
                                          "gpu_env_steps_per_sec_same_workload": gpu64["speedtest_steps_per_s"] if gpu64 else None}
        except Exception as e:
            line["config1_cpu_sanity"] = {"error": f"{type(e).__name__}: {e}"}
    # the two CPU baselines north_star names (MuJoCo C, MJX on JAX-CPU): self-reporting scripts, UNAVAILABLE in this image
    if world == 1:
        nsb = {}
        for name in ("run_mujoco_c.py", "run_mjx_cpu.py"):
            try:
                out = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", name)], capture_output=True, text=True, timeout=300).stdout.strip()
                nsb[name] = json.loads(out.splitlines()[-1]) if out else {"status": "no output"}
            except Exception as e:
                nsb[name] = {"status": "error", "why": f"{type(e).__name__}: {e}"}
        line["north_star_cpu_baselines"] = nsb
    # FP32 roofline: FLOPs counted by the oracle's stage counters on a sample of the measured states (rank 0; the counting model's formula
    # with the GPU's own statistics is kept beside it), against the MEASURED FFMA peak
    fl = None
    if not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import helpers
            from oracle import oracle as O
            O.build()
            orc = helpers.make_oracle(model, helpers.env_config())
            fl = counted_flops(orc, *flop_sample)
        except Exception as e:                                      # the counters are a reporting aid: never fail the bench line on them
            fl = {"error": f"{type(e).__name__}: {e}"}
    formula = 70e3 + mean_iters * (756.0 * mean_active + 20e3)
    flops = fl["flops_per_env_step"] if fl and "flops_per_env_step" in fl else formula
    ach_tflops = flops * n / (kernel_ms * 1e-3) / 1e12
    line["roofline_fp32"] = {"bound": "fp32 CUDA cores", "achieved": ach_tflops, "peak": ffma_peak, "unit": "TFLOP/s", "frac": ach_tflops / ffma_peak,
                             "peak_source": f"measured: FFMA microbenchmark in libmjxb.so (mjxb_ffma_peak, {ffma_ms.value:.3f} ms best of 4) -- of measured",
                             "peak_nominal": FP32_NOMINAL_TFLOPS, "frac_of_nominal": ach_tflops / FP32_NOMINAL_TFLOPS,
                             "flops_per_env_step": flops, "flops_source": "oracle stage counters (activity-aware), sample of the measured states" if fl and "flops_per_env_step" in fl else "formula 70K + n_newton (756 n_active + 20K) with the GPU's statistics",
                             "flops_formula_with_gpu_stats": formula, "counted": fl}
    print(json.dumps(line))


def sweep_point(n, v_reset, v_step, env_sys, nu, dev):
    """One batch size of BASELINE configs[0..1]: (A) speed-test semantics (cold steps in one launch, like the reference's fori_loop) and
    (B) trajectory distribution with auto-reset -- timed both as the API is called (one launch triple per step from the host) and as a
    trainer uses it (the steps replayed from a CUDA graph)."""
    import torch
    from mujoco_mjx_lab_b200 import mjx, parallel
    rank = int(os.environ.get("RANK", "0"))
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    state, obs = v_reset(torch.from_numpy(parallel.rank_keys(7, rank, n).view(np.int32)).to(dev))
    acts = [torch.randn(n, nu, device=dev, generator=g).clamp_(-1, 1) for _ in range(4)]
    rk = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device=dev, dtype=torch.int32, generator=g) for _ in range(4)]
    for i in range(64):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    nt = 200
    e0.record()
    for i in range(nt):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    e1.record()
    torch.cuda.synchronize()
    eager_ms = e0.elapsed_time(e1) / nt
    G = 32
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream())
    env_sys.reserve(n, s)
    with torch.cuda.stream(s):
        for i in range(4):
            state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        for i in range(G):
            state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    reps = 8
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / (reps * G)
    vel = torch.linspace(0, 1, n, device=dev)
    iters = 200
    mjx.speed_test(env_sys, vel, 3)
    torch.cuda.synchronize()
    e0.record()
    mjx.speed_test(env_sys, vel, iters)
    e1.record()
    torch.cuda.synchronize()
    st_ms = e0.elapsed_time(e1) / iters
    return dict(n_env=n, trajectory_graph_ms_per_step=graph_ms, trajectory_graph_steps_per_s=n / graph_ms * 1e3,
                trajectory_eager_ms_per_step=eager_ms, trajectory_eager_steps_per_s=n / eager_ms * 1e3,
                speedtest_ms_per_step=st_ms, speedtest_steps_per_s=n / st_ms * 1e3,
                note="trajectory: v_step + fused auto-reset (distribution B); graph = 32 steps per CUDA-graph replay; speedtest: "
                     "mjx_humanoid_speed_test.py semantics, 200 cold steps inside one launch (distribution A)")


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
