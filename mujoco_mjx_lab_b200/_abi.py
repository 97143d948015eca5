"""ctypes view of the C ABI structs declared in include/mjxb.h (mjxb_env_config, mjxb_state, mjxb_debug)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from .modelc import MAXU

MAXOBS = 64
AUX_DIM = 9


class EnvConfigC(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "progress_weight", "electricity_cost", "stall_torque_cost", "posture_penalty_weight", "tall_height_threshold",
        "tall_bonus_weight", "target_threshold", "target_dist", "stance_time_reward_weight", "random_joint_noise",
        "random_vel_noise", "initial_velocity_max", "terminate_height", "terminate_reward")] + \
        [(n, C.c_int32) for n in ("stop_frames", "max_episode_steps", "random_flip", "pelvis_body_id", "head_body_id",
                                   "touch_sensor_right_id", "touch_sensor_left_id", "obs_dim")] + \
        [("act_perm", C.c_int32 * MAXU), ("act_sign", C.c_float * MAXU),
         ("obs_perm", C.c_int32 * MAXOBS), ("obs_sign", C.c_float * MAXOBS)]


class StateC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("qpos", "qvel", "qacc_warmstart", "time", "aux")]


DEBUG_FIELDS = ("xpos", "xquat", "qM", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "con_dist",
                "con_pos", "con_normal", "efc_pos", "efc_D", "efc_aref", "efc_force", "efc_active", "qacc",
                "qfrc_constraint", "sensordata", "solver_niter")


class DebugC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in DEBUG_FIELDS]


def flip_permutations(cfg, nu: int, obs_dim: int):
    """act_perm/act_sign/obs_perm/obs_sign exactly as create_env_functions builds them (reference src/envs.py:49-74)."""
    act_perm, act_sign = np.arange(nu, dtype=np.int32), np.ones(nu, dtype=np.float32)
    obs_perm, obs_sign = np.arange(obs_dim, dtype=np.int32), np.ones(obs_dim, dtype=np.float32)
    if cfg.random_flip:
        def check(idx: Sequence[int], n: int, what: str):
            bad = [i for i in idx if not 0 <= i < n]
            if bad:
                raise ValueError(f"{what} indices {bad} outside [0,{n}) (the reference would silently drop them)")
        check(list(cfg.flip_action_right) + list(cfg.flip_action_left) + list(cfg.flip_action_sign), nu, "flip_action")
        check(list(cfg.flip_obs_right) + list(cfg.flip_obs_left) + list(cfg.flip_obs_sign), obs_dim, "flip_obs")
        r, l = np.array(cfg.flip_action_right, dtype=np.int64), np.array(cfg.flip_action_left, dtype=np.int64)
        act_perm[r], act_perm[l] = l, r
        act_sign[np.array(cfg.flip_action_sign, dtype=np.int64)] = -1.0
        r, l = np.array(cfg.flip_obs_right, dtype=np.int64), np.array(cfg.flip_obs_left, dtype=np.int64)
        obs_perm[r], obs_perm[l] = l, r
        obs_sign[np.array(cfg.flip_obs_sign, dtype=np.int64)] = -1.0
    return act_perm, act_sign, obs_perm, obs_sign


def make_env_config_c(cfg, nq: int, nv: int, nu: int) -> EnvConfigC:
    obs_dim = 1 + 3 + (nq - 7) + nv + 2  # reference src/envs.py:56
    if obs_dim > MAXOBS or nu > MAXU:
        raise ValueError("model exceeds compiled env capacities")
    c = EnvConfigC()
    for name, _ in EnvConfigC._fields_[:14]:
        setattr(c, name, float(getattr(cfg, name)))
    c.stop_frames, c.max_episode_steps, c.random_flip = int(cfg.stop_frames), int(cfg.max_episode_steps), int(bool(cfg.random_flip))
    for name in ("pelvis_body_id", "head_body_id", "touch_sensor_right_id", "touch_sensor_left_id"):
        v = int(getattr(cfg, name))
        if v < 0:
            raise ValueError(f"EnvConfig.{name} is unresolved; call load_model_and_create_env first")
        setattr(c, name, v)
    c.obs_dim = obs_dim
    ap, asg, op, osg = flip_permutations(cfg, nu, obs_dim)
    for i in range(nu):
        c.act_perm[i], c.act_sign[i] = int(ap[i]), float(asg[i])
    for i in range(obs_dim):
        c.obs_perm[i], c.obs_sign[i] = int(op[i]), float(osg[i])
    return c
