"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY -- see oracle/oracle.hpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
PARITY UNPINNED (no runnable mujoco-mjx here, no golden vectors in the reference): see oracle.hpp header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_DBG_FIELDS = ("xpos", "xquat", "qM", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "con_dist", "con_pos",
               "con_normal", "efc_J", "efc_pos", "efc_D", "efc_aref", "efc_force", "efc_active", "qacc", "qfrc_constraint",
               "sensordata", "solver_niter", "flops", "cdof", "cinert", "subtree_com", "qfrc_smooth", "flops_act")
_DBG_INT = {"efc_active": np.int32, "solver_niter": np.int32, "flops": np.int64, "flops_act": np.int64}


class _DebugC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _DBG_FIELDS]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (g++, no FMA contraction)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle.hpp")] + \
           [os.path.join(_HERE, "..", "include", f) for f in ("mjxb.h", "mjxb_model.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_blob_sizeof.restype = C.c_size_t
        _lib.orc_env_config_sizeof.restype = C.c_size_t
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Batched CPU evaluation of mjx.forward / mjx.step / the env layer for a compiled model blob."""

    def __init__(self, blob: np.ndarray, env_cfg_c=None, nthreads: int = 0):
        L = lib()
        if L.orc_blob_sizeof() != blob.nbytes:
            raise RuntimeError(f"blob size mismatch: C {L.orc_blob_sizeof()} vs python {blob.nbytes}")
        self.blob = np.ascontiguousarray(blob)
        self.cfg = env_cfg_c
        if env_cfg_c is not None and L.orc_env_config_sizeof() != C.sizeof(env_cfg_c):
            raise RuntimeError("env config size mismatch")
        self.nthreads = nthreads
        for k in ("nq", "nv", "nu", "nbody", "ncon", "nefc", "nsensor", "npair"):
            setattr(self, k, int(blob[k]))

    # ---- debug buffers
    def _debug(self, n: int, want) -> (Optional[_DebugC], Dict[str, np.ndarray]):
        if not want:
            return None, {}
        shapes = dict(xpos=(n, self.nbody, 3), xquat=(n, self.nbody, 4), qM=(n, self.nv, self.nv), qfrc_bias=(n, self.nv),
                      qfrc_passive=(n, self.nv), qfrc_actuator=(n, self.nv), qacc_smooth=(n, self.nv), con_dist=(n, self.ncon),
                      con_pos=(n, self.ncon, 3), con_normal=(n, self.ncon, 3), efc_J=(n, self.nefc, self.nv),
                      efc_pos=(n, self.nefc), efc_D=(n, self.nefc), efc_aref=(n, self.nefc), efc_force=(n, self.nefc),
                      efc_active=(n, self.nefc), qacc=(n, self.nv), qfrc_constraint=(n, self.nv),
                      sensordata=(n, self.nsensor), solver_niter=(n,), flops=(n,), cdof=(n, self.nv, 6),
                      cinert=(n, self.nbody, 10), subtree_com=(n, self.nbody, 3), qfrc_smooth=(n, self.nv), flops_act=(n,))
        names = _DBG_FIELDS if want is True else want
        out, dbg = {}, _DebugC()
        for name in names:
            out[name] = np.zeros(shapes[name], dtype=_DBG_INT.get(name, np.float64))
            setattr(dbg, name, out[name].ctypes.data)
        return dbg, out

    @staticmethod
    def _f64(a, shape=None):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if shape is not None:
            a = a.reshape(shape)
        return a.copy()

    # ---- mjx.forward / mjx.step
    def physics_step(self, qpos, qvel, qacc_warmstart=None, time=None, ctrl=None, nsteps: int = 1, prec: str = "f32",
                     integrate: bool = True, debug=False):
        n = np.asarray(qpos).shape[0]
        qpos, qvel = self._f64(qpos, (n, self.nq)), self._f64(qvel, (n, self.nv))
        warm = self._f64(qacc_warmstart, (n, self.nv)) if qacc_warmstart is not None else np.zeros((n, self.nv))
        time = self._f64(time, (n,)) if time is not None else np.zeros(n)
        ctrl = self._f64(ctrl, (n, self.nu)) if ctrl is not None else None
        dbg, out = self._debug(n, debug)
        lib().orc_physics_step(_p(self.blob), C.c_int(0 if prec == "f32" else 1), C.c_int(n), C.c_int(nsteps),
                               C.c_int(1 if integrate else 0), _p(qpos), _p(qvel), _p(warm), _p(time), _p(ctrl),
                               C.byref(dbg) if dbg is not None else None, C.c_int(self.nthreads))
        out.update(qpos=qpos, qvel=qvel, qacc_warmstart=warm, time=time)
        return out

    def forward(self, qpos, qvel, qacc_warmstart=None, ctrl=None, prec: str = "f32", debug=True):
        return self.physics_step(qpos, qvel, qacc_warmstart, None, ctrl, 1, prec, integrate=False, debug=debug)

    # ---- env layer
    def env_reset(self, keys, prec: str = "f32"):
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint32)).reshape(-1, 2)
        n, od = keys.shape[0], int(self.cfg.obs_dim)
        qpos, qvel, warm = np.zeros((n, self.nq)), np.zeros((n, self.nv)), np.zeros((n, self.nv))
        time, aux, obs = np.zeros(n), np.zeros((n, 9)), np.zeros((n, od))
        lib().orc_env_reset(_p(self.blob), C.byref(self.cfg), C.c_int(0 if prec == "f32" else 1), C.c_int(n), _p(keys),
                            _p(qpos), _p(qvel), _p(warm), _p(time), _p(aux), _p(obs), C.c_int(self.nthreads))
        return dict(qpos=qpos, qvel=qvel, qacc_warmstart=warm, time=time, aux=aux), obs

    def env_step(self, state: Dict[str, np.ndarray], action, prec: str = "f32", reset_keys=None, debug=False):
        n, od = np.asarray(state["qpos"]).shape[0], int(self.cfg.obs_dim)
        qpos, qvel = self._f64(state["qpos"], (n, self.nq)), self._f64(state["qvel"], (n, self.nv))
        warm, time = self._f64(state["qacc_warmstart"], (n, self.nv)), self._f64(state["time"], (n,))
        aux, action = self._f64(state["aux"], (n, 9)), self._f64(action, (n, self.nu))
        obs, reward, term, trunc = np.zeros((n, od)), np.zeros(n), np.zeros(n), np.zeros(n)
        keys = None if reset_keys is None else np.ascontiguousarray(np.asarray(reset_keys, dtype=np.uint32)).reshape(n, 2)
        mask = np.zeros(n, dtype=np.uint8)
        dbg, out = self._debug(n, debug)
        lib().orc_env_step(_p(self.blob), C.byref(self.cfg), C.c_int(0 if prec == "f32" else 1), C.c_int(n), _p(qpos), _p(qvel),
                           _p(warm), _p(time), _p(aux), _p(action), _p(obs), _p(reward), _p(term), _p(trunc), _p(keys),
                           _p(mask), C.byref(dbg) if dbg is not None else None, C.c_int(self.nthreads))
        new_state = dict(qpos=qpos, qvel=qvel, qacc_warmstart=warm, time=time, aux=aux)
        return new_state, obs, reward, term, trunc, mask, out

    def speed_test(self, vel, iters: int = 1, prec: str = "f32"):
        vel = self._f64(vel)
        pos = np.zeros_like(vel)
        lib().orc_speed_test(_p(self.blob), C.c_int(0 if prec == "f32" else 1), C.c_int(vel.shape[0]), _p(vel), _p(pos),
                             C.c_int(iters), C.c_int(self.nthreads))
        return pos


def threefry2x32(k0: int, k1: int, c0: int, c1: int):
    out = (C.c_uint32 * 2)()
    lib().orc_threefry2x32(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(c0), C.c_uint32(c1), out)
    return int(out[0]), int(out[1])


def jax_split(key, n: int) -> np.ndarray:
    key = np.ascontiguousarray(np.asarray(key, dtype=np.uint32))
    out = np.zeros((n, 2), dtype=np.uint32)
    lib().orc_jax_split(_p(key), C.c_int(n), _p(out))
    return out


def jax_uniform(key, n: int, minval: float = 0.0, maxval: float = 1.0) -> np.ndarray:
    key = np.ascontiguousarray(np.asarray(key, dtype=np.uint32))
    out = np.zeros(n, dtype=np.float32)
    lib().orc_jax_uniform(_p(key), C.c_int(n), C.c_float(minval), C.c_float(maxval), _p(out))
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())
