"""Host-side mirror of the `mujoco.mjx` calls the reference makes on its hot path, backed by libmjxb.so:

    mjx.put_model(m)        reference src/training_utils.py:105, mjx_humanoid_speed_test.py:44
    mjx.make_data(sys)      reference src/envs.py:110, mjx_humanoid_speed_test.py:51
    mjx.forward(sys, d)     reference src/envs.py:112
    mjx.step(sys, d)        reference src/envs.py:345, mjx_humanoid_speed_test.py:54

Batches are explicit (leading env axis) instead of `jax.vmap`; arrays are torch CUDA tensors whose device pointers go
straight through the C ABI on torch's current stream (no copies).  No CPU fallback: a CUDA device is required.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace as _dc_replace
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _lib, modelc
from ._abi import DEBUG_FIELDS, DebugC, EnvConfigC, StateC


class _Opt:
    def __init__(self, o: Dict[str, Any]):
        self.timestep = float(o["timestep"])
        self.iterations = int(o["iterations"])
        self.ls_iterations = int(o["ls_iterations"])
        self.tolerance = float(o["tolerance"])
        self.solver = int(o["solver"])
        self.integrator = int(o["integrator"])


class Model:
    """Stands in for `mjx.Model` (`sys` in the reference): compiled constants + a device-resident C handle."""

    def __init__(self, model: Dict[str, Any], device: Optional[torch.device] = None, env_cfg_c: Optional[EnvConfigC] = None,
                 variant: str = "fast", flags: Optional[int] = None):
        """variant: "fast" (libmjxb.so, the product) or "exact" (libmjxb_exact.so, the reference-arithmetic yardstick of the parity tests);
        flags: MJXB_FLAG_* bits for mjxb_model_create_ex (None: mjxb_model_create's defaults)."""
        if not torch.cuda.is_available():
            raise _lib.MjxbError("mujoco_mjx_lab_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.model = model
        self.blob = modelc.pack_blob(model)
        self.opt = _Opt(model["opt"])
        self.nq, self.nv, self.nu = model["nq"], model["nv"], model["nu"]
        self.nbody, self.ncon, self.nefc, self.nsensor = model["nbody"], model["ncon"], model["nefc"], model["nsensor"]
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.qpos0 = torch.tensor(np.asarray(model["qpos0"], dtype=np.float32), device=self.device)
        self.env_cfg_c = env_cfg_c
        self.variant, self.flags = variant, flags
        L = self.lib = _lib.lib(variant)
        if L.mjxb_blob_sizeof() != self.blob.nbytes or L.mjxb_env_config_sizeof() != C.sizeof(EnvConfigC):
            raise _lib.MjxbError("ABI struct size mismatch between python and libmjxb.so")
        h = C.c_void_p()
        cfg_p = C.byref(env_cfg_c) if env_cfg_c is not None else None
        if flags is None:
            _lib.check(L.mjxb_model_create(self.blob.ctypes.data_as(C.c_void_p), self.blob.nbytes, cfg_p, self.device.index or 0,
                                           C.byref(h)), "mjxb_model_create", L)
        else:
            _lib.check(L.mjxb_model_create_ex(self.blob.ctypes.data_as(C.c_void_p), self.blob.nbytes, cfg_p, self.device.index or 0,
                                              int(flags), C.byref(h)), "mjxb_model_create_ex", L)
        self.handle = h
        self.obs_dim = 1 + 3 + (self.nq - 7) + self.nv + 2

    def with_env(self, env_cfg_c: EnvConfigC) -> "Model":
        return Model(self.model, self.device, env_cfg_c, self.variant, self.flags)

    def reserve(self, n_env: int, stream: Optional[torch.cuda.Stream] = None):
        """Pre-size the launch scratch of `stream` (default: the current one) -- required before capturing it into a CUDA graph."""
        st = stream if stream is not None else torch.cuda.current_stream()
        _lib.check(self.lib.mjxb_model_reserve(self.handle, int(n_env), C.c_void_p(st.cuda_stream)), "mjxb_model_reserve", self.lib)

    def launch_config(self):
        out = (C.c_int32 * 4)()
        _lib.check(self.lib.mjxb_launch_config(self.handle, out))
        return dict(warps_per_cta=out[0], smem_bytes=out[1], num_sms=out[2], warp_smem_bytes=out[3])

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mjxb_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


@dataclass
class Data:
    """The persistent fields of `mjx.Data`: everything else is recomputed inside the step."""
    qpos: torch.Tensor            # [N, nq]
    qvel: torch.Tensor            # [N, nv]
    qacc_warmstart: torch.Tensor  # [N, nv]
    time: torch.Tensor            # [N]
    ctrl: Optional[torch.Tensor] = None  # [N, nu]

    def replace(self, **kw) -> "Data":
        return _dc_replace(self, **kw)

    def clone(self) -> "Data":
        return Data(self.qpos.clone(), self.qvel.clone(), self.qacc_warmstart.clone(), self.time.clone(),
                    None if self.ctrl is None else self.ctrl.clone())


def put_model(model: Dict[str, Any], device=None, variant: str = "fast", flags: Optional[int] = None) -> Model:
    return Model(model, device, None, variant, flags)


def make_data(sys: Model, n: int = 1) -> Data:
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=sys.device)
    return Data(sys.qpos0.unsqueeze(0).repeat(n, 1).contiguous(), z(n, sys.nv), z(n, sys.nv), z(n), z(n, sys.nu))


def _f32(t: torch.Tensor, shape) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError("expected float32 CUDA tensors")
    t = t.contiguous()
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def state_c(qpos, qvel, warm, time, aux=None) -> StateC:
    s = StateC()
    s.qpos, s.qvel, s.qacc_warmstart, s.time = qpos.data_ptr(), qvel.data_ptr(), warm.data_ptr(), time.data_ptr()
    s.aux = aux.data_ptr() if aux is not None else None
    return s


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_DBG_SHAPES = dict(xpos=lambda m: (m.nbody, 3), xquat=lambda m: (m.nbody, 4), qM=lambda m: (m.nv, m.nv), qfrc_bias=lambda m: (m.nv,),
                   qfrc_passive=lambda m: (m.nv,), qfrc_actuator=lambda m: (m.nv,), qacc_smooth=lambda m: (m.nv,),
                   con_dist=lambda m: (m.ncon,), con_pos=lambda m: (m.ncon, 3), con_normal=lambda m: (m.ncon, 3),
                   efc_pos=lambda m: (m.nefc,), efc_D=lambda m: (m.nefc,), efc_aref=lambda m: (m.nefc,), efc_force=lambda m: (m.nefc,),
                   efc_active=lambda m: (m.nefc,), qacc=lambda m: (m.nv,), qfrc_constraint=lambda m: (m.nv,),
                   sensordata=lambda m: (m.nsensor,), solver_niter=lambda m: ())


def _debug_buffers(sys: Model, n: int):
    dbg, out = DebugC(), {}
    for name in DEBUG_FIELDS:
        dt = torch.int32 if name in ("efc_active", "solver_niter") else torch.float32
        out[name] = torch.zeros((n,) + tuple(_DBG_SHAPES[name](sys)), dtype=dt, device=sys.device)
        setattr(dbg, name, out[name].data_ptr())
    return dbg, out


def _physics(sys: Model, d: Data, nsteps: int, integrate: bool, debug: bool):
    n = d.qpos.shape[0]
    qpos, qvel = _f32(d.qpos, (n, sys.nq)).clone(), _f32(d.qvel, (n, sys.nv)).clone()
    warm, time = _f32(d.qacc_warmstart, (n, sys.nv)).clone(), _f32(d.time, (n,)).clone()
    ctrl = None if d.ctrl is None else _f32(d.ctrl, (n, sys.nu))
    status = torch.zeros(n, dtype=torch.int32, device=sys.device)
    dbg, out = _debug_buffers(sys, n) if debug else (None, {})
    L = sys.lib
    st = state_c(qpos, qvel, warm, time)
    cp = C.c_void_p(ctrl.data_ptr()) if ctrl is not None else None
    dp = C.byref(dbg) if dbg is not None else None
    with torch.cuda.device(sys.device):
        if integrate:
            _lib.check(L.mjxb_physics_step(sys.handle, n, st, cp, nsteps, dp, status.data_ptr(), _stream()), "mjxb_physics_step")
        else:
            _lib.check(L.mjxb_forward(sys.handle, n, st, cp, dp, status.data_ptr(), _stream()), "mjxb_forward")
    out["status"] = status
    return Data(qpos, qvel, warm, time, d.ctrl), out


def step(sys: Model, d: Data, nsteps: int = 1, debug: bool = False):
    """mjx.step(sys, d) on a batch; `nsteps` consecutive steps stay inside one launch. Returns Data (and stage outputs if debug)."""
    nd, out = _physics(sys, d, nsteps, True, debug)
    return (nd, out) if debug else nd


def forward(sys: Model, d: Data, debug: bool = False):
    """mjx.forward(sys, d) on a batch: no integration; qacc_warmstart <- solver qacc."""
    nd, out = _physics(sys, d, 1, False, debug)
    return (nd, out) if debug else nd


def speed_test(sys: Model, vel: torch.Tensor, iters: int = 1) -> torch.Tensor:
    """mjx_humanoid_speed_test.py:48-57,88-93 device loop: `iters` x (make_data, qvel[0]=vel, step) -> qpos[0]."""
    vel = _f32(vel, (vel.shape[0],))
    pos = torch.empty_like(vel)
    with torch.cuda.device(sys.device):
        _lib.check(sys.lib.mjxb_speed_test(sys.handle, vel.shape[0], vel.data_ptr(), pos.data_ptr(), iters, _stream()), "mjxb_speed_test")
    return pos
