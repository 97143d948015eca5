"""PPO iteration driver for the B200 env path: the *caller* of the hot path (reference train_ppo.py:41-463), kept to what the
timing of BASELINE.json configs 3 and 5 needs: rollout (policy -> fused step+auto-reset), RMS observation statistics, GAE,
clipped-PPO updates with two Adam optimisers, and -- when envs are sharded over several GPUs -- NCCL all-reduce of the
gradients and of the RMS statistics.  Networks and optimiser stay in the host framework (torch here, JAX in the
reference): pure library GEMMs, not part of the kernel scope.

Same hyper-parameters and update rule as the reference:
  policy / value MLP 3x256 tanh, Glorot-normal init, state-independent log_std (reference src/networks.py:21-131)
  rollout: pre-step obs stored, reward/terminated/truncated of the step, reset obs carried (train_ppo.py:128-169)
  GAE with terminate/truncate distinction (train_ppo.py:171-202), advantage normalised per minibatch (:209)
  clipped objective + entropy bonus, separate value MSE, Adam(lr_policy) / Adam(lr_value) (:204-252)
The rollout of `rollout_length` steps is captured once in a CUDA graph (policy GEMMs, sampling, the step kernel, trajectory
stores) and replayed per iteration: the host is out of the T-loop.
"""
from __future__ import annotations

import os

import math
import time
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import parallel
from .config import PPOConfig
from .mjx import Data


def _mlp_params(in_dim: int, specs, out_dim: Optional[int], gen: torch.Generator, device):
    dims = [in_dim] + [int(h) for h, _ in specs] + ([out_dim] if out_dim is not None else [])
    params = []
    for a, b in zip(dims[:-1], dims[1:]):
        w = torch.randn(a, b, generator=gen, device=device) * math.sqrt(2.0 / (a + b))      # reference src/networks.py:44-47
        params += [w.requires_grad_(), torch.zeros(b, device=device, requires_grad=True)]
    return params


class _LinearAct(torch.autograd.Function):
    """y = tanh(x W + b) (or the plain linear layer) whose backward fuses the tanh derivative with the bias gradient in one pass
    (include/mjxb.h mjxb_tanh_bwd_colsum): torch's own path spends 70 us per layer in a column reduction of a 65536 x 256 array."""

    @staticmethod
    def forward(ctx, x, w, b, act: bool):
        y = torch.addmm(b, x, w)
        if act:
            y = torch.tanh_(y)
        ctx.save_for_backward(x, w, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes as C
        from . import _lib
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        n, c = dy.shape
        db = torch.zeros(c, dtype=dy.dtype, device=dy.device)
        dz = torch.empty_like(dy) if ctx.act else dy
        _lib.check(_lib.lib().mjxb_tanh_bwd_colsum(n, c, dy.data_ptr(), y.data_ptr() if ctx.act else None,
                                                   dz.data_ptr() if ctx.act else None, db.data_ptr(),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mjxb_tanh_bwd_colsum")
        dx = dz @ w.t() if ctx.needs_input_grad[0] else None
        # weight gradient = [k, n] x [n, c] with a tiny output and n in the tens of thousands: cuBLAS's own choice for that shape runs at
        # ~90 TFLOP/s (tools/wgrad_probe.py: 100-118 us); an explicit 64-way split over the rows as one batched GEMM + sum takes 43 us
        split = 64
        if n >= 8192 and n % split == 0:
            dw = torch.bmm(x.view(split, n // split, x.shape[1]).transpose(1, 2), dz.view(split, n // split, c)).sum(0)
        else:
            dw = x.t() @ dz
        return dx, dw, db, None


def _mlp_apply(params, x, n_hidden: int):
    fused_bwd = x.is_cuda and torch.is_grad_enabled() and params[0].requires_grad and x.dtype == torch.float32
    for i in range(0, len(params), 2):
        act = i // 2 < n_hidden
        if fused_bwd:
            x = _LinearAct.apply(x, params[i], params[i + 1], act)
        else:
            x = torch.addmm(params[i + 1], x, params[i])
            if act:
                x = torch.tanh(x)
    return x


def gaussian_logprob(mean, log_std, action):  # reference train_ppo.py:121-126
    var = torch.exp(2.0 * log_std)
    return -0.5 * torch.sum((action - mean) ** 2 / var + 2.0 * log_std + math.log(2.0 * math.pi), dim=-1)


class _PPOLoss(torch.autograd.Function):
    """Policy loss of one minibatch (reference train_ppo.py:204-232) with its gradients produced by the forward launches
    (include/mjxb.h mjxb_ppo_loss): replaces ~45 elementwise / reduction launches of the autograd chain by two."""

    @staticmethod
    def forward(ctx, mean, log_std, action, old_logp, adv, clip_eps: float, ent_coef: float):
        """`mean` may be wider than the action (an output layer padded to a multiple of four columns, see PPOTrainer): its first
        action.shape[1] columns are the mean, the rest is ignored and receives a zero gradient."""
        import ctypes as C
        from . import _lib
        mean, action, old_logp, adv = mean.contiguous(), action.contiguous(), old_logp.contiguous(), adv.contiguous()
        n, ld = mean.shape
        a = action.shape[1]
        g_mean, g_ls = torch.empty_like(mean), torch.empty_like(log_std)
        scratch = torch.empty(5, dtype=torch.float32, device=mean.device)          # [0:4] statistics, [4] the loss
        _lib.check(_lib.lib().mjxb_ppo_loss_ld(n, a, ld, mean.data_ptr(), log_std.data_ptr(), action.data_ptr(), old_logp.data_ptr(),
                                               adv.data_ptr(), float(clip_eps), float(ent_coef), scratch.data_ptr(), g_mean.data_ptr(),
                                               g_ls.data_ptr(), scratch[4:].data_ptr(),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mjxb_ppo_loss_ld")
        ctx.save_for_backward(g_mean, g_ls)
        return scratch[4]

    @staticmethod
    def backward(ctx, g):
        g_mean, g_ls = ctx.saved_tensors
        return g_mean * g, g_ls * g, None, None, None, None, None


class _FlatAdam:
    """optax.adam / torch.optim.Adam over ONE flat parameter / gradient buffer (include/mjxb.h mjxb_adam): two launches per step, the
    step counter lives on the device (CUDA-graph replayable); [0, split) uses lr0, the rest lr1."""

    def __init__(self, flat_p, flat_g, split: int, lr0: float, lr1: float, eps: float = 1e-8, comm=None):
        self.p, self.g, self.split, self.lr0, self.lr1, self.eps = flat_p, flat_g, split, lr0, lr1, eps
        self.m, self.v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=flat_p.device)
        self.comm = comm                                                   # _PeerComm: gradients of all ranks summed inside the Adam kernel

    def step(self, grad_scale: float = 1.0):
        import ctypes as C
        from . import _lib
        if self.comm is not None:      # one kernel: cross-GPU barrier, sum of every rank's gradients over NVLink peer memory, Adam, barrier
            _lib.check(_lib.lib().mjxb_allreduce_adam(self.comm.handle, self.p.numel(), self.split, self.p.data_ptr(), self.m.data_ptr(),
                                                      self.v.data_ptr(), self.step_dev.data_ptr(), self.lr0, self.lr1, 0.9, 0.999, self.eps,
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mjxb_allreduce_adam")
            return
        _lib.check(_lib.lib().mjxb_adam(self.p.numel(), self.split, self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                        self.step_dev.data_ptr(), self.lr0, self.lr1, 0.9, 0.999, self.eps, float(grad_scale),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mjxb_adam")


class _PeerComm:
    """Gradient buffers of all ranks mapped into each other's address space (CUDA IPC over NVLink / NVSwitch, include/mjxb.h mjxb_comm_*):
    the learner's all-reduce then happens INSIDE the Adam kernel (mjxb_allreduce_adam), not as an NCCL launch."""

    def __init__(self, n_floats: int, device):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        self.L, self.n = L, n_floats
        rank, world = dist.get_rank(), dist.get_world_size()
        # every step that can fail locally is followed by a collective agreement, so that either all ranks use the communicator or none
        h, blob, ok = C.c_void_p(), (C.c_char * 128)(), True
        try:
            _lib.check(L.mjxb_comm_create(rank, world, n_floats, C.byref(h)), "mjxb_comm_create")
            _lib.check(L.mjxb_comm_local_handles(h, blob), "mjxb_comm_local_handles")
        except Exception as e:
            ok, self._why = False, str(e)
        self.handle = h if ok else None
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, bytes(blob.raw)))
        if not all(g[0] for g in gathered):
            raise RuntimeError("CUDA IPC export failed on a rank")
        allb = b"".join(g[1] for g in gathered)
        ok = L.mjxb_comm_connect(h, C.c_char_p(allb)) == 0
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        if not all(flags):
            raise RuntimeError("CUDA IPC import (peer access) failed on a rank: " + L.mjxb_last_cuda_error().decode())
        ptr = L.mjxb_comm_grad_buffer(h)

        class _Ext:   # zero-copy torch view of the library-owned (IPC-exported) gradient buffer
            __cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        self._ext = _Ext()
        self.grad = torch.as_tensor(self._ext, device=device)
        assert self.grad.data_ptr() == int(ptr)
        dist.barrier()

    def error(self) -> int:
        return int(self.L.mjxb_comm_error(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.L.mjxb_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


@dataclass
class RMS:  # reference src/training_utils.py:20-56
    mean: torch.Tensor
    var: torch.Tensor
    count: torch.Tensor

    @classmethod
    def create(cls, dim, device):
        return cls(torch.zeros(dim, device=device), torch.ones(dim, device=device), torch.tensor(1e-4, device=device))

    def update(self, x: torch.Tensor, world: int):
        """reference src/training_utils.py:33-52 (update_rms): Chan merge of the batch moments into the running ones, variance floored at
        1e-4. The batch moments are centred and accumulated in float64 (a one-pass E[x^2]-E[x]^2 in float32 cancels catastrophically on a
        near-constant observation and can go negative); with sharded envs the per-rank (n, mean, M2) are Chan-merged across ranks through
        one all-reduce of (n, n*mean, M2 + n*mean^2) in float64."""
        xd = x.double()
        n = torch.tensor(float(x.shape[0]), device=x.device, dtype=torch.float64)
        b_mean = xd.mean(0)
        m2 = ((xd - b_mean) ** 2).sum(0)
        if world > 1:
            packed = torch.cat([b_mean * n, m2 + n * b_mean * b_mean, n.reshape(1)])
            dist.all_reduce(packed)
            d = x.shape[1]
            n = packed[-1]
            b_mean = packed[:d] / n
            m2 = torch.clamp_min(packed[d: 2 * d] - n * b_mean * b_mean, 0.0)
        b_var = m2 / n
        mean, var, count = self.mean.double(), self.var.double(), self.count.double()
        delta = b_mean - mean
        tot = count + n
        new_mean = mean + delta * n / tot
        new_var = (var * count + b_var * n + delta * delta * count * n / tot) / tot
        self.mean = new_mean.float()
        self.var = torch.clamp_min(new_var, 1e-4).float()      # reference: new_var = max(new_var, 1e-4)
        self.count = tot.float()

    def normalize(self, x):
        return torch.clamp((x - self.mean) / torch.sqrt(self.var + 1e-8), -10.0, 10.0)


class PPOTrainer:
    def __init__(self, cfg: PPOConfig, v_reset, v_step, num_envs_local: int, seed: int = 42, use_cuda_graph: bool = True,
                 use_fused_policy: bool = True, use_fused_learner: bool = True):
        self.cfg, self.v_reset, self.v_step = cfg, v_reset, v_step
        self.sys = v_step.sys
        self.dev = self.sys.device
        self.n, self.T = num_envs_local, cfg.rollout_length
        self.rank, _, self.world = parallel.dist_env()
        if self.world > 1 and not dist.is_initialized():
            parallel.init("nccl")
        od, nu = self.sys.obs_dim, self.sys.nu
        gen = torch.Generator(device=self.dev).manual_seed(seed)           # same init on every rank
        self.nh_p, self.nh_v = len(cfg.policy_hidden_layer_specs), len(cfg.value_hidden_layer_specs)
        self.policy = _mlp_params(od, cfg.policy_hidden_layer_specs, nu, gen, self.dev)
        self.log_std = torch.full((nu,), float(cfg.log_std_init), device=self.dev, requires_grad=True)
        self.value = _mlp_params(od, cfg.value_hidden_layer_specs, 1, gen, self.dev)
        self.fused_learner = bool(use_fused_learner) and self.dev.type == "cuda" and nu <= 32
        self.kpad = od
        if self.fused_learner:
            # every parameter is a view into ONE flat buffer and its .grad a view into one flat gradient buffer: autograd accumulates
            # in place, the NCCL all-reduce runs on the flat gradient without a flatten / unflatten pass, Adam is one kernel.
            # The learner's copies of the input layers and of the policy's output layer are zero-padded (54 -> 64 input rows, 21 -> 32
            # output columns): a float32 GEMM whose leading dimension is not a multiple of four falls back to cuBLAS's unaligned
            # mma.sync kernels (tools/prof_update2.py: 0.58 ms of a 1.37 ms minibatch step); padded, every GEMM of the step runs on the
            # tcgen05 TF32 kernels. Padding rows / columns see zero inputs / receive zero gradients, so they stay zero under Adam:
            # the function computed is the unpadded network's (self.policy / self.value are views of the unpadded blocks).
            def _pad(t, shape):
                out = torch.zeros(shape, device=t.device)
                out[tuple(slice(0, d) for d in t.shape)] = t.detach()
                return out
            self.kpad = ((od + 15) // 16) * 16 if od % 4 else od
            npad = ((nu + 15) // 16) * 16 if nu % 4 else nu
            hp, hv = self.policy[0].shape[1], self.value[0].shape[1]
            pol_pad = [_pad(self.policy[0], (self.kpad, hp))] + self.policy[1:-2] + [_pad(self.policy[-2], (self.policy[-2].shape[0], npad)),
                                                                                    _pad(self.policy[-1], (npad,))]
            val_pad = [_pad(self.value[0], (self.kpad, hv))] + self.value[1:]
            allp = pol_pad + [self.log_std] + val_pad
            # every parameter starts on a 256-byte boundary of the flat buffer (gaps are zeros with zero gradients): a weight whose
            # POINTER is not 16-byte aligned sends its GEMMs to the same unaligned kernels as an unaligned leading dimension
            # (the 21-float log_std used to misalign every value-network weight behind it)
            offs, o = [], 0
            for p in allp:
                offs.append(o)
                o += ((p.numel() + 63) // 64) * 64
            n_pol = offs[len(pol_pad) + 1]
            self.flat_p = torch.zeros(o, device=self.dev)
            for p, o_ in zip(allp, offs):
                self.flat_p[o_:o_ + p.numel()] = p.detach().reshape(-1)
            self.comm = None
            if self.world > 1 and os.environ.get("MJXB_PPO_NCCL_ALLREDUCE") is None:
                try:                                                    # peer-memory gradient exchange fused into the Adam kernel
                    self.comm = _PeerComm(self.flat_p.numel(), self.dev)
                except Exception as e:                                  # no CUDA IPC / peer access: NCCL all-reduce inside the update graph
                    print(f"[ppo] peer-memory communicator unavailable ({type(e).__name__}: {e}); using NCCL")
                    self.comm = None
            self.flat_g = self.comm.grad if self.comm is not None else torch.zeros_like(self.flat_p)
            views = []
            for p, o in zip(allp, offs):
                v = self.flat_p[o:o + p.numel()].view(p.shape).detach().requires_grad_()
                v.grad = self.flat_g[o:o + p.numel()].view(p.shape)
                views.append(v)
            npol = len(self.policy)
            self._pol_pad, self.log_std, self._val_pad = views[:npol], views[npol], views[npol + 1:]
            # the unpadded blocks, for everything outside the minibatch step (rollout fallback, tests, checkpoints)
            self.policy = [self._pol_pad[0].detach()[:od]] + [p.detach() for p in self._pol_pad[1:-2]] + \
                          [self._pol_pad[-2].detach()[:, :nu], self._pol_pad[-1].detach()[:nu]]
            self.value = [self._val_pad[0].detach()[:od]] + [p.detach() for p in self._val_pad[1:]]
            self.opt = _FlatAdam(self.flat_p, self.flat_g, n_pol, cfg.lr_policy, cfg.lr_value, comm=self.comm)
            self.opt_p = self.opt_v = None
        else:
            self.opt_p = torch.optim.Adam(self.policy + [self.log_std], lr=cfg.lr_policy, eps=1e-8, capturable=True)
            self.opt_v = torch.optim.Adam(self.value, lr=cfg.lr_value, eps=1e-8, capturable=True)
        self.rms = RMS.create(od, self.dev)
        self.gen = torch.Generator(device=self.dev).manual_seed(seed + 1000 * (self.rank + 1))
        torch.cuda.manual_seed(seed + 7919 * (self.rank + 1))
        keys = torch.from_numpy(parallel.rank_keys(seed, self.rank, self.n).view(np.int32)).to(self.dev)
        (d, aux), obs = v_reset(keys)
        self.state = (Data(d.qpos, d.qvel, d.qacc_warmstart, d.time), aux)
        f32 = dict(dtype=torch.float32, device=self.dev)
        T, n = self.T, self.n
        # observations of a rollout live in one [T + 1, n, obs_dim] buffer: the step writes slot t + 1 directly (no per-step copy);
        # `obs` (the current observation) is the last slot, `obs_traj` the first T
        self.obs_all = torch.empty(T + 1, n, od, **f32)
        self.obs_traj, self.obs = self.obs_all[:T], self.obs_all[T]
        self.obs.copy_(obs)
        self.act_traj = torch.empty(T, n, nu, **f32)
        self.logp_traj, self.r_traj = torch.empty(T, n, **f32), torch.empty(T, n, **f32)
        self.term_traj, self.trunc_traj = torch.empty(T, n, **f32), torch.empty(T, n, **f32)
        self.fused = None
        if use_fused_policy and self.dev.type == "cuda":
            from . import policy as _policy
            flat = [p.detach() for p in (self._pol_pad if self.fused_learner else self.policy)]   # (zero-padded or not: same image)
            if _policy.supported(flat, od, nu) and all(a == "tanh" for _, a in cfg.policy_hidden_layer_specs):
                self.fused = _policy.FusedPolicy(flat, self.log_std.detach(), od, nu)
        self.graph = None
        self.upd = None
        self.use_graph = use_cuda_graph
        self.n_grads = self.flat_p.numel() if self.fused_learner else sum(p.numel() for p in self.policy + [self.log_std] + self.value)
        self.timing: Dict[str, float] = {}

    # ---------------------------------------------------------------- rollout (train_ppo.py:128-169)
    def _rollout_body(self):
        if self.fused is not None:
            self.fused.pack()                                           # bf16 image of the current policy weights (4 tiny kernels)
        nu = self.act_traj.shape[-1]
        # reset keys (and, when they fit in 512 MB, the sampling noise) of the whole rollout are drawn by one launch each
        keys_all = torch.randint(-2 ** 31, 2 ** 31 - 1, (self.T, self.n, 2), device=self.dev, dtype=torch.int32)
        eps_all = torch.randn(self.T, self.n, nu, device=self.dev) if self.T * self.n * nu * 4 <= (512 << 20) else None
        self.obs_all[0].copy_(self.obs)
        for t in range(self.T):
            keys = keys_all[t]
            obs_t = self.obs_all[t]
            obs_next = self.obs if t == self.T - 1 else self.obs_all[t + 1]
            if self.fused is not None:
                # normalise -> MLP -> sample -> log-prob in one tcgen05 launch, written straight into the trajectory buffers
                eps = eps_all[t] if eps_all is not None else torch.randn(self.n, nu, device=self.dev)
                act, _ = self.fused.act(obs_t, eps, self.rms.mean, self.rms.var, act_out=self.act_traj[t], logp_out=self.logp_traj[t])
            else:
                obs_n = self.rms.normalize(obs_t)
                mean = _mlp_apply(self.policy, obs_n, self.nh_p)
                eps = torch.randn(mean.shape, device=self.dev)          # default CUDA generator: graph-capture safe
                act = mean + torch.exp(self.log_std) * eps
                self.act_traj[t].copy_(act)
                self.logp_traj[t].copy_(gaussian_logprob(mean, self.log_std, act))
            # the step writes the next observation and this step's reward / terminated / truncated where the trainer keeps them
            self.v_step.autoreset(self.state, act, keys, inplace=True,
                                  out=(obs_next, self.r_traj[t], self.term_traj[t], self.trunc_traj[t]))

    @torch.no_grad()
    def collect_rollout(self):
        if not self.use_graph:
            self._rollout_body()
            return
        if self.graph is None:
            # warm-up outside capture (allocator pools, lazy library state, the overflow list), then capture the whole T-loop
            s = self.cap_stream = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream())
            self.sys.reserve(self.n, s)                                 # the step's launch scratch is per stream: size it before capturing
            with torch.cuda.stream(s):
                saved = [t.clone() for t in (self.state[0].qpos, self.state[0].qvel, self.state[0].qacc_warmstart, self.state[0].time,
                                             self.state[1], self.obs)]
                T_full, self.T = self.T, min(2, self.T)
                self._rollout_body()
                self.T = T_full
                for dst, src in zip((self.state[0].qpos, self.state[0].qvel, self.state[0].qacc_warmstart, self.state[0].time,
                                     self.state[1], self.obs), saved):
                    dst.copy_(src)
            torch.cuda.current_stream().wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            self.rms_mean_buf, self.rms_var_buf = self.rms.mean.clone(), self.rms.var.clone()
            graph_rms = RMS(self.rms_mean_buf, self.rms_var_buf, self.rms.count)
            real_rms, self.rms = self.rms, graph_rms
            with torch.cuda.graph(self.graph, stream=s):             # capture on the stream whose scratch was reserved
                self._rollout_body()
            self.rms = real_rms
        self.rms_mean_buf.copy_(self.rms.mean)
        self.rms_var_buf.copy_(self.rms.var)
        self.graph.replay()

    def check_health(self):
        """Once per rollout (one host sync): a tensor-core completion that was never observed by the fused policy kernel, or non-finite
        rewards / observations from the step, must stop training instead of being learned from."""
        if self.fused is not None and int(self.fused.error) != 0:
            raise RuntimeError("fused policy kernel reported a tensor-core completion timeout")
        if getattr(self, "comm", None) is not None and self.comm.error() != 0:
            raise RuntimeError("a peer did not reach the gradient barrier within the bounded wait (mjxb_allreduce_adam)")
        if not (bool(torch.isfinite(self.r_traj).all()) and bool(torch.isfinite(self.obs).all())):
            raise RuntimeError("non-finite reward / observation in the rollout (MJXB_STATUS_NAN)")

    # ---------------------------------------------------------------- GAE (train_ppo.py:171-202)
    @torch.no_grad()
    def compute_gae(self, rewards, values, terminated, truncated, force_torch: bool = False):
        g, lam = self.cfg.gamma, self.cfg.lam
        if rewards.is_cuda and not force_torch:                      # one launch: include/mjxb.h mjxb_gae
            import ctypes as C
            from . import _lib
            values = values.contiguous()
            adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
            _lib.check(_lib.lib().mjxb_gae(rewards.shape[0], rewards.shape[1], rewards.data_ptr(), values.data_ptr(), terminated.data_ptr(),
                                           truncated.data_ptr(), float(g), float(lam), adv.data_ptr(), ret.data_ptr(),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mjxb_gae")
            return adv, ret
        delta = rewards + g * values[1:] * (1.0 - terminated) - values[:-1]          # [T, n] in three kernels
        decay = g * lam * (1.0 - torch.maximum(terminated, truncated))
        adv = torch.empty_like(rewards)
        carry = torch.zeros_like(rewards[0])
        for t in range(rewards.shape[0] - 1, -1, -1):                                 # reverse scan: one fused addcmul per step
            carry = torch.addcmul(delta[t], decay[t], carry)
            adv[t] = carry
        return adv, adv + values[:-1]

    def _zero_grads(self):
        if self.fused_learner:
            self.flat_g.zero_()
        else:
            self.opt_p.zero_grad(set_to_none=True)
            self.opt_v.zero_grad(set_to_none=True)

    def _opt_step(self):
        """Adam on both networks; sharded runs divide the all-reduced gradient sum by the world size first."""
        if self.fused_learner:
            self.opt.step(1.0 / self.world)
        else:
            self.opt_p.step()
            self.opt_v.step()

    def _minibatch_fb(self, obs_f, act_f, logp_f, ret_f, adv_f, idx, zero: bool):
        """Losses of one minibatch (train_ppo.py:204-252) and their gradients."""
        cfg = self.cfg
        o, a, olp, r_, ad = obs_f[idx], act_f[idx], logp_f[idx], ret_f[idx], adv_f[idx]
        if zero:
            self._zero_grads()
        mean = _mlp_apply(self._pol_pad if self.fused_learner else self.policy, o, self.nh_p)
        if self.fused_learner:
            loss = _PPOLoss.apply(mean, self.log_std, a, olp, ad, cfg.clip_eps, cfg.ent_coef)       # (mean: act_dim columns + padding)
        else:
            logp = gaussian_logprob(mean, self.log_std, a)
            ratio = torch.exp(logp - olp)
            ad_n = (ad - ad.mean()) / (ad.std(unbiased=False) + 1e-8)
            loss_p = -torch.minimum(ratio * ad_n, torch.clamp(ratio, 1 - cfg.clip_eps, 1 + cfg.clip_eps) * ad_n).mean()
            entropy = 0.5 * torch.sum(1.0 + math.log(2.0 * math.pi) + 2.0 * self.log_std) / a.shape[-1]
            loss = loss_p - cfg.ent_coef * entropy
        loss.backward()
        v = _mlp_apply(self._val_pad if self.fused_learner else self.value, o, self.nh_v).squeeze(-1)
        loss_v = torch.mean((v - r_) ** 2)
        loss_v.backward()

    def _capture_update(self):
        """One CUDA graph per minibatch step: gather + forward + backward + (sharded runs) the NCCL all-reduce of the flat gradient,
        captured INSIDE the graph, + Adam. A minibatch is then a single replay with no host round trip between the backward pass and
        the optimiser (the eager all-reduce wedged between two replays cost ~185 us per minibatch at 8 GPUs, profiles/r1_bench_n8.json).
        If this NCCL build refuses stream capture the collective stays eager between two graphs (u["st"] is then the second half)."""
        u = self.upd
        params = self.policy + [self.log_std] + self.value
        if self.fused_learner:
            u["flat"] = self.flat_g
        torch.cuda.synchronize()

        def unflatten_and_step():
            if self.fused_learner:                                      # the collective ran on the flat gradient buffer itself
                self._opt_step()
                return
            o = 0
            for p in params:
                p.grad.copy_(u["flat"][o:o + p.numel()].view_as(p.grad) / self.world)
                o += p.numel()
            self._opt_step()

        def capture(fused_collective: bool):
            if not self.fused_learner:
                self._zero_grads()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._minibatch_fb(u["obs"], u["act"], u["logp"], u["ret"], u["adv"], u["idx"], zero=self.fused_learner)
                if self.world == 1 or getattr(self, "comm", None) is not None:
                    self._opt_step()          # sharded + peer communicator: the collective is inside this kernel
                else:
                    if not self.fused_learner:
                        u["flat"].copy_(torch.cat([p.grad.reshape(-1) for p in params]))
                    if fused_collective:
                        dist.all_reduce(u["flat"])
                        unflatten_and_step()
            return g

        u["st"] = None
        peer = getattr(self, "comm", None) is not None
        fused = self.world > 1 and not peer and os.environ.get("MJXB_PPO_EAGER_ALLREDUCE") is None
        if fused:
            try:
                u["fb"] = capture(True)
            except Exception as e:                                      # capture of the collective unsupported: keep it eager
                print(f"[ppo] NCCL all-reduce could not be captured ({type(e).__name__}: {e}); using an eager all-reduce between two graphs")
                torch.cuda.synchronize()
                fused = False
        if not fused:
            u["fb"] = capture(False)
            if self.world > 1 and not peer:
                u["st"] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(u["st"], pool=u["fb"].pool()):
                    unflatten_and_step()
        u["collective_in_graph"] = bool(fused or peer)
        u["collective"] = "peer-memory sum fused into the Adam kernel" if peer else ("NCCL all-reduce" if self.world > 1 else "none")
        # capture does not execute: run this minibatch now
        u["fb"].replay()
        if u["st"] is not None:
            dist.all_reduce(u["flat"])
            u["st"].replay()

    def _allreduce_grads(self, params):
        if self.world == 1 or getattr(self, "comm", None) is not None:      # (peer communicator: summed inside the Adam kernel)
            return
        if self.fused_learner:                 # sum over ranks; _opt_step divides by the world size
            dist.all_reduce(self.flat_g)
            return
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(flat)
        flat /= self.world
        o = 0
        for p in params:
            p.grad.copy_(flat[o:o + p.numel()].view_as(p.grad))
            o += p.numel()

    # ---------------------------------------------------------------- one PPO iteration (train_ppo.py:321-371)
    def iteration(self) -> Dict[str, float]:
        cfg, T, n = self.cfg, self.T, self.n
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        self.collect_rollout()
        ev[1].record()
        od = self.obs_traj.shape[-1]
        with torch.no_grad():
            self.rms.update(self.obs_traj.reshape(-1, od), self.world)
            if self.kpad != od:                 # normalised observations in rows of kpad floats (zero padded) for the aligned GEMMs
                if getattr(self, "_obs_pad", None) is None:
                    self._obs_pad = torch.zeros(T + 1, n, self.kpad, device=self.dev)
                self._obs_pad[:, :, :od] = self.rms.normalize(self.obs_all)
                obs_norm, stack = self._obs_pad[:T], self._obs_pad.view((T + 1) * n, self.kpad)
                values = _mlp_apply(self._val_pad, stack, self.nh_v).reshape(T + 1, n)
            else:
                obs_norm = self.rms.normalize(self.obs_traj)
                obs_last = self.rms.normalize(self.obs)
                stack = torch.cat([obs_norm, obs_last.unsqueeze(0)], 0).reshape((T + 1) * n, od)
                values = _mlp_apply(self.value, stack, self.nh_v).reshape(T + 1, n)
            adv, ret = self.compute_gae(self.r_traj, values, self.term_traj, self.trunc_traj)
        obs_f, act_f = obs_norm.reshape(T * n, self.kpad), self.act_traj.reshape(T * n, -1)
        logp_f, adv_f, ret_f = self.logp_traj.reshape(-1), adv.reshape(-1), ret.reshape(-1)
        total = T * n
        mb = min(cfg.minibatch_size, total)
        steps_per_epoch = total // mb
        if self.use_graph:
            # the minibatch step (gather, both networks forward/backward, Adam) replays from CUDA graphs over static buffers: the
            # eager step is ~200 launches and is bound by their dispatch, not by the GPU
            if self.upd is None:
                f32 = dict(dtype=torch.float32, device=self.dev)
                self.upd = dict(obs=torch.empty(total, self.kpad, **f32), adv=torch.empty(total, **f32), ret=torch.empty(total, **f32),
                                idx=torch.zeros(mb, dtype=torch.int64, device=self.dev), act=act_f, logp=logp_f, eager_steps=0,
                                fb=None, st=None, flat=torch.zeros(self.n_grads, **f32))
            u = self.upd
            u["obs"].copy_(obs_f); u["adv"].copy_(adv_f); u["ret"].copy_(ret_f)
        for _ in range(cfg.epochs):
            perm = torch.randperm(total, device=self.dev, generator=self.gen)[: steps_per_epoch * mb].view(steps_per_epoch, mb)
            for idx in perm:
                if not self.use_graph:
                    self._minibatch_fb(obs_f, act_f, logp_f, ret_f, adv_f, idx, zero=True)
                    self._allreduce_grads(self.policy + [self.log_std] + self.value)
                    self._opt_step()
                    continue
                u = self.upd
                u["idx"].copy_(idx)
                if u["eager_steps"] < 3:        # warm-up on real minibatches (cuBLAS workspaces, Adam state) before the capture
                    self._minibatch_fb(u["obs"], u["act"], u["logp"], u["ret"], u["adv"], u["idx"], zero=True)
                    self._allreduce_grads(self.policy + [self.log_std] + self.value)
                    self._opt_step()
                    u["eager_steps"] += 1
                    continue
                if u["fb"] is None:
                    self._capture_update()      # captures, then runs this minibatch
                    continue
                u["fb"].replay()
                if u["st"] is not None:
                    dist.all_reduce(u["flat"])
                    u["st"].replay()
        ev[2].record()
        torch.cuda.synchronize()
        self.check_health()                                             # after the timed region: the iteration already synchronised
        done = torch.maximum(self.term_traj, self.trunc_traj).sum()
        out = {"rollout_ms": ev[0].elapsed_time(ev[1]), "update_ms": ev[1].elapsed_time(ev[2]), "iter_ms": ev[0].elapsed_time(ev[2]),
               "train_return_avg": float(self.r_traj.sum(0).mean()), "train_eplen_avg": float(total / max(float(done), 1.0)),
               "minibatches": cfg.epochs * steps_per_epoch, "allreduce_floats": self.n_grads if self.world > 1 else 0,
               "collective_in_graph": bool(self.upd and self.upd.get("collective_in_graph")),
               "collective": (self.upd or {}).get("collective", "none")}
        return out


def time_ppo(num_envs_local: int, rollout_length: int, iters: int = 5, warmup: int = 3, minibatch_size: int = 65536,
             use_cuda_graph: bool = True, model=None, env_cfg=None) -> Dict[str, float]:
    """Times PPO iterations (device-synchronised, max over ranks). Returns the mean over `iters` timed iterations."""
    from . import modelc, training_utils
    from .config import EnvConfig
    cfg = PPOConfig()
    cfg.rollout_length, cfg.minibatch_size, cfg.epochs = rollout_length, minibatch_size, 4         # reference src/config.json:117-120
    cfg.lr_value, cfg.gamma = 3e-4, 0.99
    cfg.env_config = env_cfg or EnvConfig(posture_penalty_weight=0.0, random_flip=True)
    model = model or modelc.builtin_model("humanoid_mjx")
    _, _, _, _, _, _, _, _, v_reset, v_step = training_utils.load_model_and_create_env("", cfg.env_config, model=model)
    # JAX's default matmul precision on NVIDIA GPUs is TF32 for float32 operands (reference runs with jax defaults): same here
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    tr = PPOTrainer(cfg, v_reset, v_step, num_envs_local, use_cuda_graph=use_cuda_graph)
    for _ in range(warmup):
        tr.iteration()
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    acc = {"rollout_ms": 0.0, "update_ms": 0.0, "iter_ms": 0.0}
    last = {}
    for _ in range(iters):
        last = tr.iteration()
        for k in acc:
            acc[k] += last[k]
    torch.cuda.synchronize()
    wall = parallel.max_over_ranks((time.perf_counter() - t0) / iters * 1e3, tr.dev)
    res = {k: parallel.max_over_ranks(v / iters, tr.dev) for k, v in acc.items()}
    res.update(wall_iter_ms=wall, envs_per_gpu=num_envs_local, rollout_length=rollout_length, world=tr.world,
               env_steps_per_sec=num_envs_local * tr.world * rollout_length / (wall * 1e-3), minibatches=last.get("minibatches"),
               allreduce_floats_per_minibatch=last.get("allreduce_floats"), collective_in_graph=last.get("collective_in_graph"),
               collective=last.get("collective"), train_return_avg=last.get("train_return_avg"),
               cuda_graph=bool(use_cuda_graph), fused_policy_kernel=tr.fused is not None)
    return res
