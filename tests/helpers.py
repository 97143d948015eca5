"""Shared test helpers: model/oracle construction and seeded state generators."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mujoco_mjx_lab_b200 import _abi, config, jax_random, modelc  # noqa: E402


def env_config(**kw):
    """EnvConfig with the effective PPO defaults (reference src/config.json overlay: posture 0, random_flip true)."""
    cfg = config.EnvConfig(posture_penalty_weight=0.0, random_flip=True)
    for k, v in kw.items():
        setattr(cfg, k, v)
    cfg.pelvis_body_id, cfg.head_body_id, cfg.touch_sensor_right_id, cfg.touch_sensor_left_id = 4, 2, 0, 1
    return cfg


def load(name="humanoid_mjx", overrides=None):
    m = modelc.builtin_model(name)
    if overrides:
        m["opt"] = dict(m["opt"], **overrides)
    return m


def make_oracle(model, cfg=None, nthreads=0):
    from oracle import oracle as O
    cc = _abi.make_env_config_c(cfg, model["nq"], model["nv"], model["nu"]) if cfg is not None else None
    return O.Oracle(modelc.pack_blob(model), cc, nthreads)


def rand_quat(rng, n, max_angle):
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rng.uniform(-max_angle, max_angle, size=(n, 1))
    return np.concatenate([np.cos(ang / 2), ax * np.sin(ang / 2)], axis=1)


def make_states(model, n, seed, kind):
    """Returns qpos[n,nq], qvel[n,nv], warm[n,nv], ctrl[n,nu] (float32-representable float64 arrays)."""
    rng = np.random.default_rng(seed)
    nq, nv, nu = model["nq"], model["nv"], model["nu"]
    q = np.tile(model["qpos0"], (n, 1))
    v = np.zeros((n, nv))
    if kind == "stand":          # reset-like noise, feet pressed 0..4 mm into the floor
        q[:, 7:] += rng.uniform(-0.01, 0.01, (n, nq - 7))
        q[:, 2] -= rng.uniform(0.0005, 0.004, n)
        v = rng.uniform(-0.05, 0.05, (n, nv))
    elif kind == "free":         # lifted clear of the floor, arbitrary joint angles inside the limits
        lo, hi = model["jnt_range"][1:, 0], model["jnt_range"][1:, 1]
        q[:, 7:] = rng.uniform(lo * 0.8, hi * 0.8, (n, nq - 7))
        q[:, 2] += rng.uniform(1.0, 2.0, n)
        q[:, 3:7] = rand_quat(rng, n, np.pi)
        v = rng.normal(size=(n, nv)) * 1.0
    elif kind == "tumble":       # random poses near the floor: many contacts and limit rows, big velocities
        lo, hi = model["jnt_range"][1:, 0], model["jnt_range"][1:, 1]
        q[:, 7:] = rng.uniform(lo * 1.05, hi * 1.05, (n, nq - 7))
        q[:, 2] = rng.uniform(0.25, 1.2, n)
        q[:, 3:7] = rand_quat(rng, n, 1.2)
        v = rng.normal(size=(n, nv)) * 2.0
    elif kind == "lean":         # standing poses leaning / crouching: feet contacts with varied foot tilt
        q[:, 7:] += rng.uniform(-0.25, 0.25, (n, nq - 7))
        q[:, 3:7] = rand_quat(rng, n, 0.25)
        q[:, 2] -= rng.uniform(0.0, 0.08, n)
        v = rng.normal(size=(n, nv)) * 0.5
    elif kind == "crumple":      # folded far beyond the joint limits and sunk into the floor: > 32 candidate contacts, ~100 rows
        q[:, 7:] = rng.uniform(-3.14, 3.14, (n, nq - 7))
        q[:, 2] = rng.uniform(-0.6, -0.2, n)
        q[:, 3:7] = rand_quat(rng, n, 3.0)
        v = rng.normal(size=(n, nv))
    else:
        raise ValueError(kind)
    warm = rng.normal(size=(n, nv)) * (0.0 if kind == "free" else 5.0)
    ctrl = np.clip(rng.normal(size=(n, nu)), -1, 1)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    return f32(q), f32(v), f32(warm), f32(ctrl)


def ppo_keys(seed, n):
    """Per-env reset keys the way train_ppo.py:117-118 derives them."""
    rng = jax_random.PRNGKey(seed)
    _, k = jax_random.split(rng, 2)
    return jax_random.split(k, n)


# ---------------------------------------------------------------- host build of the reverse-mode step (tests/adjoint_host.cpp)
_ADJ_LIB = None


def adjoint_host():
    """ctypes handle of tests/_build/libadjoint_host.so: csrc/mjxb_adjoint.cuh compiled for the CPU in double (test harness only)."""
    global _ADJ_LIB
    if _ADJ_LIB is not None:
        return _ADJ_LIB
    import ctypes as C
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_build", "libadjoint_host.so")
    srcs = [os.path.join(here, "adjoint_host.cpp")] + [os.path.join(ROOT, "mujoco_mjx_lab_b200", "csrc", f) for f in ("mjxb_adjoint.cuh", "mjxb_model_dev.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-o", out, srcs[0]],
                       check=True, capture_output=True)
    _ADJ_LIB = C.CDLL(out)
    return _ADJ_LIB


def tangent_perturb(model, q, d, eps):
    """qpos moved by eps along tangent coordinate d (dof index): hinge / free translation add, free rotation q (x) exp(eps e_k / 2)."""
    q = np.array(q, dtype=np.float64)
    if d < 3:
        q[d] += eps
    elif d < 6:
        k = d - 3
        r = np.zeros(4); r[0] = np.cos(eps / 2); r[1 + k] = np.sin(eps / 2)
        w, x, y, z = q[3:7] / np.linalg.norm(q[3:7])
        a = np.array([w, x, y, z])
        q[3:7] = np.array([a[0] * r[0] - a[1] * r[1] - a[2] * r[2] - a[3] * r[3],
                           a[0] * r[1] + a[1] * r[0] + a[2] * r[3] - a[3] * r[2],
                           a[0] * r[2] - a[1] * r[3] + a[2] * r[0] + a[3] * r[1],
                           a[0] * r[3] + a[1] * r[2] - a[2] * r[1] + a[3] * r[0]])
    else:
        q[d + 1] += eps
    return q
