"""Drop-in mirror of the reference env factory (reference src/envs.py:26-497) on top of libmjxb.so.

    single_reset, single_step, v_reset, v_step = create_env_functions(sys, cfg, q0, nq, nv)
    (d, aux), obs = v_reset(keys)                                   # keys: uint32[N, 2] threefry key data
    (d, aux), obs, reward, terminated, truncated = v_step((d, aux), action)

Same names, argument meaning and return structure as the reference; arrays are torch CUDA tensors with a leading env
axis (the reference's `jax.vmap`).  `state[0]` exposes `.qpos` (N, 28) / `.qvel` (N, 27) as the reference's rendering
and APG code expect (reference src/rendering.py:160-164).  Extra (not in the reference): `v_step.autoreset(state, action,
keys)` fuses the trainer's reset-and-merge glue (reference train_ppo.py:143-161) into the same launch.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Tuple

import numpy as np
import torch

from . import _lib
from ._abi import AUX_DIM, make_env_config_c
from .mjx import Data, Model, _f32, _stream, state_c

EnvState = Tuple[Data, torch.Tensor]


def _keys_tensor(keys, device) -> torch.Tensor:
    """uint32[N,2] key data as an int32-bit-pattern CUDA tensor."""
    if isinstance(keys, np.ndarray):
        keys = torch.from_numpy(np.ascontiguousarray(keys.astype(np.uint32)).view(np.int32))
    if keys.dtype == torch.uint32:
        keys = keys.view(torch.int32)
    if keys.dtype == torch.int64:
        keys = (keys & 0xFFFFFFFF).to(torch.int64)
        keys = torch.where(keys >= 2 ** 31, keys - 2 ** 32, keys).to(torch.int32)
    if keys.dtype != torch.int32:
        raise TypeError("keys must be uint32 / int32 / int64 key data of shape [N, 2]")
    return keys.to(device).contiguous()


def create_env_functions(sys: Model, cfg, q0, nq: int, nv: int) -> Tuple[Callable, Callable, Callable, Callable]:
    if nq != sys.nq or nv != sys.nv:
        raise ValueError("nq/nv do not match the compiled model")
    q0 = np.asarray(q0, dtype=np.float32)
    if not np.array_equal(q0, np.asarray(sys.model["qpos0"], dtype=np.float32)):
        raise ValueError("q0 must be the model's qpos0 (reference src/training_utils.py:107)")
    env_sys = sys.with_env(make_env_config_c(cfg, nq, nv, sys.nu))
    L = env_sys.lib
    dev, nu, od = env_sys.device, env_sys.nu, env_sys.obs_dim
    f32 = dict(dtype=torch.float32, device=dev)

    def _alloc_state(n):
        d = Data(torch.empty(n, nq, **f32), torch.empty(n, nv, **f32), torch.empty(n, nv, **f32), torch.empty(n, **f32))
        return d, torch.empty(n, AUX_DIM, **f32)

    def v_reset(keys):
        k = _keys_tensor(keys, dev)
        if k.dim() != 2 or k.shape[1] != 2:
            raise ValueError("keys must have shape [N, 2]")
        n = k.shape[0]
        d, aux = _alloc_state(n)
        obs = torch.empty(n, od, **f32)
        with torch.cuda.device(dev):
            _lib.check(L.mjxb_reset(env_sys.handle, n, k.data_ptr(), state_c(d.qpos, d.qvel, d.qacc_warmstart, d.time, aux),
                                    obs.data_ptr(), None, _stream()), "mjxb_reset")
        return (d, aux), obs

    def _step(state: EnvState, action, keys=None, inplace=False, out=None):
        d, aux = state
        n = d.qpos.shape[0]
        action = _f32(action, (n, nu))
        qpos, qvel = _f32(d.qpos, (n, nq)), _f32(d.qvel, (n, nv))
        warm, time, aux = _f32(d.qacc_warmstart, (n, nv)), _f32(d.time, (n,)), _f32(aux, (n, AUX_DIM))
        if inplace:
            d2, aux2 = Data(qpos, qvel, warm, time), aux
        else:
            d2, aux2 = _alloc_state(n)
        if out is not None and not all(t.is_contiguous() for t in out):
            raise ValueError("output buffers must be contiguous")
        if out is not None:         # caller-owned output buffers (obs [n, obs_dim], reward / terminated / truncated [n]), e.g. rollout slices
            obs, reward, term, trunc = (_f32(out[0], (n, od)), _f32(out[1], (n,)), _f32(out[2], (n,)), _f32(out[3], (n,)))
        else:
            obs = torch.empty(n, od, **f32)
            reward, term, trunc = torch.empty(n, **f32), torch.empty(n, **f32), torch.empty(n, **f32)
        sin, sout = state_c(qpos, qvel, warm, time, aux), state_c(d2.qpos, d2.qvel, d2.qacc_warmstart, d2.time, aux2)
        with torch.cuda.device(dev):
            if keys is None:
                _lib.check(L.mjxb_step(env_sys.handle, n, sin, action.data_ptr(), sout, obs.data_ptr(), reward.data_ptr(),
                                       term.data_ptr(), trunc.data_ptr(), None, _stream()), "mjxb_step")
            else:
                k = _keys_tensor(keys, dev)
                _lib.check(L.mjxb_step_autoreset(env_sys.handle, n, sin, action.data_ptr(), k.data_ptr(), sout, obs.data_ptr(),
                                                 reward.data_ptr(), term.data_ptr(), trunc.data_ptr(), None, None, _stream()),
                           "mjxb_step_autoreset")
        return (d2, aux2), obs, reward, term, trunc

    def v_step(state: EnvState, action):
        return _step(state, action)

    v_step.autoreset = lambda state, action, keys, inplace=False, out=None: _step(state, action, keys, inplace, out)
    v_step.sys = env_sys

    def single_reset(key):
        (d, aux), obs = v_reset(np.asarray(key, dtype=np.uint32).reshape(1, 2) if not torch.is_tensor(key) else key.reshape(1, 2))
        return (Data(d.qpos[0], d.qvel[0], d.qacc_warmstart[0], d.time[0]), aux[0]), obs[0]

    def single_step(state: EnvState, action):
        d, aux = state
        db = Data(d.qpos.reshape(1, nq), d.qvel.reshape(1, nv), d.qacc_warmstart.reshape(1, nv), d.time.reshape(1))
        (d2, aux2), obs, r, te, tr = v_step((db, aux.reshape(1, AUX_DIM)), action.reshape(1, nu))
        return (Data(d2.qpos[0], d2.qvel[0], d2.qacc_warmstart[0], d2.time[0]), aux2[0]), obs[0], r[0], te[0], tr[0]

    return single_reset, single_step, v_reset, v_step
