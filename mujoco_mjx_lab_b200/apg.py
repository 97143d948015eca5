"""Analytic policy gradients through the B200-native env step (BASELINE.json configs[3]; reference train_apg.py:96-209).

The reference differentiates `rollout_return` with `jax.value_and_grad` through `lax.scan(jax.checkpoint(v_step))`
(train_apg.py:161-209): every step is re-run forward during the backward sweep and then transposed by XLA.  Here the step is
`DiffStep`, a `torch.autograd.Function` whose forward is `mjxb_step_fwd_tape` and whose backward is the hand-written reverse-mode
kernel `mjxb_step_vjp` (include/mjxb.h, csrc/mjxb_adjoint.cuh): it keeps the step's inputs plus the solver's qacc (108 B / env-step)
and recomputes the rest, as `jax.checkpoint` does.

The reference's own `train_apg.py` is stale against its env API: it passes `cfg.lighten_solver` where
`load_model_and_create_env` expects the EnvConfig, and unpacks four values from `v_step` which returns five.  `APGTrainer` is the
corrected caller: same solver settings (CG, 4 iterations, 4 line-search iterations, train_apg.py:101-105), same policy
(tanh-squashed MLP on [qpos, qvel], src/networks.py:63-80), same discounted-return loss with `done` cutting the discount
(train_apg.py:170-181), optax.chain(clip_by_global_norm(0.3), adam(lr)) as clip + Adam (train_apg.py:139-143).
NCCL appears only in the gradient all-reduce when envs are sharded over ranks.
"""
from __future__ import annotations

import ctypes as C
import math
import time
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, parallel
from ._abi import AUX_DIM
from .config import APGConfig, EnvConfig
from .mjx import Data, _stream, state_c


class DiffStep(torch.autograd.Function):
    """(qpos, qvel, qacc_warmstart, time, aux, action) -> (qpos', qvel', qacc_warmstart', time', aux', obs, reward, terminated, truncated),
    differentiable in qpos, qvel, aux[1,2,3,7] and action; the cotangents of qpos', qvel', aux' and reward are pulled back."""

    @staticmethod
    def forward(ctx, sysm, qpos, qvel, warm, tm, aux, action):
        n = qpos.shape[0]
        f32 = dict(dtype=torch.float32, device=qpos.device)
        qpos, qvel, warm, tm, aux, action = (t.detach().contiguous() for t in (qpos, qvel, warm, tm, aux, action))
        qpos2, qvel2, warm2 = torch.empty_like(qpos), torch.empty_like(qvel), torch.empty_like(warm)
        tm2, aux2 = torch.empty_like(tm), torch.empty_like(aux)
        obs = torch.empty(n, sysm.obs_dim, **f32)
        reward, term, trunc = torch.empty(n, **f32), torch.empty(n, **f32), torch.empty(n, **f32)
        with torch.cuda.device(qpos.device):
            _lib.check(sysm.lib.mjxb_step_fwd_tape(sysm.handle, n, state_c(qpos, qvel, warm, tm, aux), action.data_ptr(),
                                                   state_c(qpos2, qvel2, warm2, tm2, aux2), obs.data_ptr(), reward.data_ptr(), term.data_ptr(),
                                                   trunc.data_ptr(), warm2.data_ptr(), None, _stream()), "mjxb_step_fwd_tape", sysm.lib)
        ctx.sysm = sysm
        ctx.save_for_backward(qpos, qvel, warm, tm, aux, action, warm2)     # warm2 = the solver's qacc = the tape
        ctx.mark_non_differentiable(warm2, tm2, obs, term, trunc)
        return qpos2, qvel2, warm2, tm2, aux2, obs, reward, term, trunc

    @staticmethod
    def backward(ctx, g_qpos2, g_qvel2, _gw, _gt, g_aux2, _go, g_reward, _gte, _gtr):
        qpos, qvel, warm, tm, aux, action, tape = ctx.saved_tensors
        sysm, n = ctx.sysm, qpos.shape[0]
        ptr = lambda t: None if t is None else t.contiguous().data_ptr()
        keep = [None if t is None else t.contiguous() for t in (g_qpos2, g_qvel2, g_aux2, g_reward)]
        g_qpos, g_qvel, g_aux, g_act = torch.empty_like(qpos), torch.empty_like(qvel), torch.empty_like(aux), torch.empty_like(action)
        with torch.cuda.device(qpos.device):
            _lib.check(sysm.lib.mjxb_step_vjp(sysm.handle, n, state_c(qpos, qvel, warm, tm, aux), action.data_ptr(), tape.data_ptr(),
                                              ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), g_qpos.data_ptr(), g_qvel.data_ptr(),
                                              g_aux.data_ptr(), g_act.data_ptr(), None, _stream()), "mjxb_step_vjp", sysm.lib)
        return None, g_qpos, g_qvel, None, None, g_aux, g_act


def diff_step(sysm, state, action):
    """Differentiable `v_step` (no auto-reset): ((d, aux), action) -> ((d', aux'), obs, reward, terminated, truncated)."""
    d, aux = state
    qpos2, qvel2, warm2, tm2, aux2, obs, reward, term, trunc = DiffStep.apply(sysm, d.qpos, d.qvel, d.qacc_warmstart, d.time, aux, action)
    return (Data(qpos2, qvel2, warm2, tm2), aux2), obs, reward, term, trunc


def _policy_params(in_dim, hidden, depth, out_dim, gen, device, output_scale: float = 1.0):
    dims = [in_dim] + [hidden] * depth + [out_dim]
    params = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        w = torch.randn(a, b, generator=gen, device=device) * math.sqrt(2.0 / (a + b))      # reference src/networks.py:44-47
        if i == len(dims) - 2:
            w = w * output_scale
        params += [w.requires_grad_(), torch.zeros(b, device=device, requires_grad=True)]
    return params


def _policy_apply(params, x):
    """APGPolicy.apply (reference src/networks.py:78-80): tanh MLP body, linear head, tanh squashing."""
    n = len(params) // 2
    for i in range(n):
        x = torch.addmm(params[2 * i + 1], x, params[2 * i])
        if i < n - 1:
            x = torch.tanh(x)
    return torch.tanh(x)


class APGTrainer:
    def __init__(self, cfg: APGConfig, v_reset, v_step, batch_size_local: int, env_cfg=None, seed: int = 0, output_scale: float = 1.0,
                 use_cuda_graph: bool = True):
        self.cfg, self.v_reset, self.v_step = cfg, v_reset, v_step
        self.sys = v_step.sys
        self.dev = self.sys.device
        self.n, self.H = batch_size_local, cfg.horizon
        self.rank, _, self.world = parallel.dist_env()
        if self.world > 1 and not dist.is_initialized():
            parallel.init("nccl")
        self.obs_dim = self.sys.nq + self.sys.nv                                  # reference train_apg.py:119
        gen = torch.Generator(device=self.dev).manual_seed(seed)
        self.params = _policy_params(self.obs_dim, cfg.hidden_size, cfg.hidden_depth, self.sys.nu, gen, self.dev, output_scale)
        self.opt = torch.optim.Adam(self.params, lr=cfg.lr, eps=1e-8, capturable=True)
        self.use_graph, self.graph, self.eager_steps = bool(use_cuda_graph), None, 0
        self.keys_buf = torch.zeros(self.n, 2, dtype=torch.int32, device=self.dev)
        self.norm_flag = torch.zeros((), dtype=torch.bool, device=self.dev)       # train_apg.py:171-175 jnp.where(use_norm, ...)
        self.obs_mean = torch.zeros(self.obs_dim, device=self.dev)
        self.obs_var = torch.ones(self.obs_dim, device=self.dev)
        self.obs_count = 1e-4
        self.step_no = 0
        self.seed = seed

    def rollout_return(self, keys, use_norm=None):
        """reference train_apg.py:161-190. Returns (mean discounted return, obs trajectory, mean reward)."""
        state, _ = self.v_reset(keys)
        n = self.n
        disc = torch.ones(n, device=self.dev)
        acc = torch.zeros(n, device=self.dev)
        obs_traj, r_sum = [], 0.0
        for _ in range(self.H):
            d, aux = state
            obs = torch.cat([d.qpos, d.qvel], dim=1)
            x = torch.where(self.norm_flag, torch.clamp((obs - self.obs_mean) / (torch.sqrt(self.obs_var) + 1e-8), -10.0, 10.0), obs)
            act = _policy_apply(self.params, torch.nan_to_num(x, nan=0.0, posinf=1e6, neginf=-1e6))
            state, _, r, te, tr = diff_step(self.sys, state, act)
            done = torch.maximum(te, tr)
            # Guards the reference does not have (its trainer stops on a non-finite loss, train_apg.py:278): with the 4-iteration CG of
            # train_apg.py:101-105 a few fallen bodies per thousand diverge numerically within ~100 steps (the CPU restatement of mjx.step
            # does the same, tools/apg_stability_probe.py). An env that has already finished contributes nothing (0 * nan must stay 0),
            # and an env whose reward is non-finite or absurd is treated as finished at that step.
            diverged = (~torch.isfinite(r)) | (r.abs() > 1e3)
            r = torch.where((disc > 0) & ~diverged, r, torch.zeros_like(r))
            done = torch.maximum(done, diverged.float())
            acc = acc + disc * r
            disc = disc * self.cfg.gamma * (1.0 - done)
            obs_traj.append(obs.detach())
            r_sum = r_sum + r.detach().mean()
        return acc.mean(), torch.stack(obs_traj), r_sum / self.H

    def _fwd_bwd_step(self, ev=None):
        """One optimisation step on self.keys_buf (everything on the current stream, capturable): value_and_grad of -return through the
        rollout, gradient all-reduce when sharded, clip_by_global_norm(0.3), Adam. Returns loss / statistics as device tensors."""
        for p in self.params:
            if p.grad is not None:
                p.grad.zero_()
        if ev:
            ev[0].record()
        ret, obs_traj, mean_reward = self.rollout_return(self.keys_buf)
        loss = -ret
        if ev:
            ev[1].record()
        loss.backward()
        if self.world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in self.params])
            dist.all_reduce(flat)
            flat /= self.world
            o = 0
            for p in self.params:
                p.grad.copy_(flat[o:o + p.numel()].view_as(p.grad))
                o += p.numel()
        grad_norm = torch.nn.utils.clip_grad_norm_(self.params, 0.3, foreach=True)          # optax.clip_by_global_norm(0.3)
        self.opt.step()
        if ev:
            ev[2].record()
        return loss.detach(), mean_reward, grad_norm.detach(), obs_traj

    def update(self) -> Dict[str, float]:
        """reference train_apg.py:197-209 + :253-262: one optimisation step. After two eager steps the whole update -- 128 policy /
        step launches forward, the reverse sweep through mjxb_step_vjp, clip, Adam -- is captured once and replayed as ONE CUDA graph
        (the eager loop is bound by ~2500 launch dispatches per update, not by the GPU)."""
        keys = torch.from_numpy(parallel.rank_keys(self.seed + 1 + self.step_no, self.rank, self.n).view(np.int32)).to(self.dev)
        self.keys_buf.copy_(keys)
        self.norm_flag.fill_(bool(self.cfg.normalize_observations and self.step_no >= 100))   # warm-up without normalisation (:238)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not self.use_graph or self.eager_steps < 2:
            t0.record()
            self._out = self._fwd_bwd_step(ev)
            t1.record()
            self.eager_steps += 1
            split = True
        else:
            if self.graph is None:
                s = torch.cuda.Stream(device=self.dev)
                s.wait_stream(torch.cuda.current_stream())
                self.sys.reserve(self.n, s)                              # launch scratch of the capture stream (include/mjxb.h)
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=s):
                    self._out = self._fwd_bwd_step(None)
            t0.record()
            self.graph.replay()
            t1.record()
            split = False
        loss, mean_reward, grad_norm, obs_traj = self._out
        if self.cfg.normalize_observations and self.step_no % 10 == 0:                      # :289-291 (in place: the graph reads these)
            x = obs_traj.reshape(-1, self.obs_dim).double()
            x = torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)
            bm, bv, bn = x.mean(0), x.var(0, unbiased=False), float(x.shape[0])
            delta, tot = bm - self.obs_mean.double(), self.obs_count + bn
            mean = self.obs_mean.double() + delta * bn / tot
            m2 = self.obs_var.double() * self.obs_count + bv * bn + delta * delta * self.obs_count * bn / tot
            self.obs_mean.copy_(mean.float())
            self.obs_var.copy_(torch.clamp_min(m2 / tot, 1e-4).float())
            self.obs_count = tot
        self.step_no += 1
        torch.cuda.synchronize()
        out = {"loss": float(loss), "mean_reward": float(mean_reward), "grad_norm": float(grad_norm), "update_ms": t0.elapsed_time(t1),
               "cuda_graph": not split}
        if split:
            out["forward_ms"], out["backward_ms"] = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        if not math.isfinite(out["loss"]):
            raise RuntimeError("non-finite APG loss (reference train_apg.py:278-287 stops here as well)")
        return out


def make_apg_env(model=None, env_cfg=None):
    """The env the way train_apg.py:101-112 builds it: lighten_solver, then CG with 4 iterations and 4 line-search iterations."""
    from . import modelc, training_utils
    cfg = APGConfig()
    env_cfg = env_cfg or EnvConfig()                                                          # dataclass defaults (APG does not load config.json)
    model = model or modelc.builtin_model("humanoid_mjx")
    solver_options = {"solver": 1, "iterations": 4, "ls_iterations": 4}                       # mujoco.mjtSolver.mjSOL_CG = 1
    out = training_utils.load_model_and_create_env("", env_cfg, lighten_solver=cfg.lighten_solver, solver_options=solver_options, model=model)
    return cfg, out


def time_apg(batch_size_local: int, horizon: int, iters: int = 3, warmup: int = 2, hidden_size: int = 32,
             output_scale: float = 0.01) -> Dict[str, float]:
    """Times APG updates (forward rollout + reverse sweep + optimiser), device-synchronised, max over ranks.
    `output_scale` shrinks the random-init action head: with the reference's Xavier-scale head, velocity feedback through 40-120 N m
    gears makes the CG-4/4 physics of train_apg.py:101-105 diverge to non-finite states within ~30 of the 128 steps (the CPU restatement
    of mjx.step does the same, tools/apg_stability_probe.py), and the reference's trainer stops on a non-finite loss (train_apg.py:278)."""
    cfg, env = make_apg_env()
    cfg.horizon, cfg.hidden_size = horizon, hidden_size
    tr = APGTrainer(cfg, env[8], env[9], batch_size_local, output_scale=output_scale)
    eager = {}
    for _ in range(max(warmup, 3)):                                   # two eager steps (timed forward / backward split), then the capture
        o = tr.update()
        if "forward_ms" in o:
            eager = {"eager_forward_ms": o["forward_ms"], "eager_backward_ms": o["backward_ms"], "eager_update_ms": o["update_ms"]}
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc = {"update_ms": 0.0}
    last = {}
    for _ in range(iters):
        last = tr.update()
        for k in acc:
            acc[k] += last[k]
    torch.cuda.synchronize()
    wall = parallel.max_over_ranks((time.perf_counter() - t0) / iters * 1e3, tr.dev)
    res = {k: parallel.max_over_ranks(v / iters, tr.dev) for k, v in acc.items()}
    res.update(eager)
    res["cuda_graph"] = bool(last.get("cuda_graph"))
    res.update(wall_update_ms=wall, envs_per_gpu=batch_size_local, horizon=horizon, world=tr.world, solver="CG 4/4 (train_apg.py:101-105)",
               env_steps_per_sec=batch_size_local * tr.world * horizon / (wall * 1e-3), loss=last.get("loss"), grad_norm=last.get("grad_norm"),
               policy_output_scale=output_scale,
               reverse_mode="mjxb_step_vjp (implicit-function adjoint of the solve, hand-written kernels)")
    return res
