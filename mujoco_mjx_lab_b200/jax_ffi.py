"""JAX binding of libmjxb.so through XLA FFI custom calls (csrc/mjxb_ffi.cc) -- the adapter BASELINE.json's north_star names.

With it the reference keeps JAX for its networks and optax and only swaps the env:

    from mujoco_mjx_lab_b200.jax_ffi import create_env_functions      # instead of src.envs.create_env_functions
    single_reset, single_step, v_reset, v_step = create_env_functions(sys, cfg, q0, nq, nv)

`v_reset` / `v_step` are callable under `jax.jit` / `lax.scan` (reference train_ppo.py:143,166-168), state leaves keep a leading env
axis (`state[0].qpos` (N, 28), reference src/rendering.py:160-164), `v_step` is differentiable through `jax.custom_vjp` whose backward
is the hand-written reverse-mode kernel (reference train_apg.py:161-209), and `v_step.autoreset(state, action, keys)` is the fused
reset-and-merge of train_ppo.py:143-161.  Buffers are zero-copy (XLA device buffers in, state aliased in place).

NOT TESTED HERE: jax / jaxlib are absent from this image and there is no network (tests/test_jax_ffi.py skips with that reason).
Importing this module without jax raises ImportError; nothing else in the package imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import NamedTuple

try:
    import jax
    import jax.numpy as jnp
except ImportError as e:  # pragma: no cover - this image
    raise ImportError("mujoco_mjx_lab_b200.jax_ffi needs jax with CUDA support (pip install jax[cuda12]==0.7.2); "
                      "the torch host mirror mujoco_mjx_lab_b200.envs needs no JAX") from e

import numpy as np

from . import _lib, modelc
from ._abi import AUX_DIM, make_env_config_c

_HERE = os.path.dirname(os.path.abspath(__file__))
_FFI_PATH = os.path.join(_HERE, "libmjxb_ffi.so")
_registered = False


class Data(NamedTuple):
    """The persistent fields of mjx.Data (a pytree: jit / scan / tree_map friendly)."""
    qpos: jax.Array
    qvel: jax.Array
    qacc_warmstart: jax.Array
    time: jax.Array


def build_ffi() -> str:
    if not os.path.exists(_FFI_PATH):
        subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "ffi"], check=True)
    return _FFI_PATH


def _register():
    global _registered
    if _registered:
        return
    _lib.lib()                                             # libmjxb.so first: libmjxb_ffi.so links against it
    ffi = C.CDLL(build_ffi())
    for name, sym in (("mjxb_reset", "MjxbReset"), ("mjxb_step", "MjxbStep"), ("mjxb_step_autoreset", "MjxbStepAutoreset"),
                      ("mjxb_step_vjp", "MjxbStepVjp")):
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(ffi, sym)), platform="CUDA")
    _registered = True


def create_env_functions(sys, cfg, q0, nq: int, nv: int):
    """Same signature and return 4-tuple as reference src/envs.py:26,497. `sys` is the compiled-constants dict of
    mujoco_mjx_lab_b200.modelc (stands in for mjx.Model)."""
    _register()
    L = _lib.lib()
    blob = modelc.pack_blob(sys)
    nu = int(sys["nu"])
    od = 1 + 3 + (nq - 7) + nv + 2
    cfg_c = make_env_config_c(cfg, nq, nv, nu)
    handle = C.c_void_p()
    _lib.check(L.mjxb_model_create(blob.ctypes.data_as(C.c_void_p), blob.nbytes, C.byref(cfg_c), jax.devices("gpu")[0].id, C.byref(handle)),
               "mjxb_model_create")
    h = np.int64(handle.value)
    f32 = lambda *s: jax.ShapeDtypeStruct(s, jnp.float32)
    state_types = lambda n: (f32(n, nq), f32(n, nv), f32(n, nv), f32(n), f32(n, AUX_DIM))
    step_types = lambda n: state_types(n) + (f32(n, od), f32(n), f32(n), f32(n))
    alias = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4}

    def v_reset(keys):
        kd = jax.random.key_data(keys) if jnp.issubdtype(keys.dtype, jax.dtypes.prng_key) else keys
        n = kd.shape[0]
        qpos, qvel, warm, time, aux, obs = jax.ffi.ffi_call("mjxb_reset", state_types(n) + (f32(n, od),))(kd.astype(jnp.uint32), model_handle=h)
        return (Data(qpos, qvel, warm, time), aux), obs

    def _step_raw(d, aux, action):
        n = d.qpos.shape[0]
        return jax.ffi.ffi_call("mjxb_step", step_types(n))(d.qpos, d.qvel, d.qacc_warmstart, d.time, aux, action, model_handle=h)

    @jax.custom_vjp
    def _step(d, aux, action):
        qpos, qvel, warm, time, aux2, obs, r, te, tr = _step_raw(d, aux, action)
        return (Data(qpos, qvel, warm, time), aux2), obs, r, te, tr

    def _step_fwd(d, aux, action):
        out = _step(d, aux, action)
        return out, (d, aux, action, out[0][0].qacc_warmstart)     # tape: inputs + the solver's qacc

    def _step_bwd(res, g):
        d, aux, action, tape = res
        (gd, g_aux), _g_obs, g_r, _, _ = g
        n = d.qpos.shape[0]
        g_qpos, g_qvel, g_aux_in, g_act = jax.ffi.ffi_call("mjxb_step_vjp", (f32(n, nq), f32(n, nv), f32(n, AUX_DIM), f32(n, nu)))(
            d.qpos, d.qvel, d.qacc_warmstart, d.time, aux, action, tape, gd.qpos, gd.qvel, g_aux, g_r, model_handle=h)
        zeros = jnp.zeros_like
        return Data(g_qpos, g_qvel, zeros(d.qacc_warmstart), zeros(d.time)), g_aux_in, g_act

    _step.defvjp(_step_fwd, _step_bwd)

    def v_step(state, action):
        d, aux = state
        return _step(d, aux, action)

    def autoreset(state, action, keys):
        d, aux = state
        n = d.qpos.shape[0]
        kd = jax.random.key_data(keys) if jnp.issubdtype(keys.dtype, jax.dtypes.prng_key) else keys
        qpos, qvel, warm, time, aux2, obs, r, te, tr = jax.ffi.ffi_call("mjxb_step_autoreset", step_types(n), input_output_aliases=alias)(
            d.qpos, d.qvel, d.qacc_warmstart, d.time, aux, action, kd.astype(jnp.uint32), model_handle=h)
        return (Data(qpos, qvel, warm, time), aux2), obs, r, te, tr

    v_step.autoreset = autoreset
    v_step.handle = handle

    def single_reset(key):
        (d, aux), obs = v_reset(jnp.reshape(jax.random.key_data(key) if jnp.issubdtype(key.dtype, jax.dtypes.prng_key) else key, (1, 2)))
        return (jax.tree_util.tree_map(lambda x: x[0], d), aux[0]), obs[0]

    def single_step(state, action):
        d, aux = state
        (d2, aux2), obs, r, te, tr = v_step((jax.tree_util.tree_map(lambda x: x[None], d), aux[None]), action[None])
        return (jax.tree_util.tree_map(lambda x: x[0], d2), aux2[0]), obs[0], r[0], te[0], tr[0]

    return single_reset, single_step, jax.jit(v_reset), v_step
