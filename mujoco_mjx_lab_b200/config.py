"""Host-side mirror of the reference's config dataclasses (reference src/config.py:8-172).

Field names, defaults and the JSON overlay rules are the reference's, so a `src/config.json` written for the
reference loads unchanged.  One deliberate deviation: flip index lists that point outside
the action/observation vector raise ValueError instead of being silently dropped.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import List, Tuple


def _ilist(*v):
    return field(default_factory=lambda: list(v))


@dataclass
class BaseConfig:  # reference src/config.py:8-26
    xml_path: str = "models/humanoid_mjx.xml"
    lighten_solver: bool = False
    seed: int = 42
    checkpoint_every: int = 50
    log_interval: int = 10
    eval_interval: int = 50
    results_dir: str = "results"
    save_video: bool = True
    render_fps: int = 60
    render_duration: float = 6.0
    camera_name: str = "side_view"


@dataclass
class EnvConfig:  # reference src/config.py:29-66
    progress_weight: float = 1.0
    electricity_cost: float = 0.026
    stall_torque_cost: float = 0.0000023
    joints_at_limit_cost: float = 5.0
    posture_penalty_weight: float = 0.60
    tall_height_threshold: float = 0.7
    tall_bonus_weight: float = 0.0
    target_threshold: float = 0.15
    target_dist: float = 2.0
    stop_frames: int = 1
    stance_time_reward_weight: float = 0.0
    random_joint_noise: float = 0.01
    random_vel_noise: float = 0.01
    initial_velocity_max: float = 0.5
    terminate_height: float = 0.7
    terminate_reward: float = 0.0
    max_episode_steps: int = 1000
    random_flip: bool = False
    joint_limit_force_threshold: float = 6.5
    # ids resolved by load_model_and_create_env
    pelvis_body_id: int = -1
    head_body_id: int = -1
    touch_sensor_right_id: int = -1
    touch_sensor_left_id: int = -1
    # left/right symmetry index lists
    flip_action_right: List[int] = _ilist(3, 4, 5, 6, 7, 8, 15, 16, 17)
    flip_action_left: List[int] = _ilist(9, 10, 11, 12, 13, 14, 18, 19, 20)
    flip_action_sign: List[int] = _ilist(0, 2)
    flip_obs_right: List[int] = _ilist(7, 8, 9, 10, 11, 12, 19, 20, 21, 34, 35, 36, 37, 38, 39, 46, 47, 48)
    flip_obs_left: List[int] = _ilist(13, 14, 15, 16, 17, 18, 22, 23, 24, 40, 41, 42, 43, 44, 45, 49, 50, 51)
    flip_obs_sign: List[int] = _ilist(1, 3, 4, 6, 26, 28, 30, 31, 33, 52)


@dataclass
class APGConfig(BaseConfig):  # reference src/config.py:69-86
    lighten_solver: bool = True
    hidden_size: int = 32
    hidden_depth: int = 2
    batch_size: int = 8
    horizon: int = 24
    gamma: float = 0.99
    lr: float = 5e-5
    total_steps: int = 8000
    normalize_observations: bool = True


@dataclass
class PPOConfig(BaseConfig):  # reference src/config.py:89-172
    lighten_solver: bool = False
    env_config: EnvConfig = field(default_factory=EnvConfig)
    policy_hidden_layer_specs: List[Tuple[int, str]] = field(default_factory=lambda: [(256, "tanh")] * 3)
    value_hidden_layer_specs: List[Tuple[int, str]] = field(default_factory=lambda: [(256, "tanh")] * 3)
    num_envs: int = 2048
    rollout_length: int = 128
    gamma: float = 0.999
    lam: float = 0.95
    lr_policy: float = 3e-4
    lr_value: float = 1e-3
    clip_eps: float = 0.2
    ent_coef: float = 0.01
    vf_coef: float = 0.5
    epochs: int = 4
    minibatch_size: int = 1024
    log_std_init: float = 0.0
    total_iterations: int = 1000

    @property
    def total_steps(self) -> int:
        return self.total_iterations

    @classmethod
    def from_json(cls, path: str) -> "PPOConfig":
        cfg = cls()
        if not os.path.exists(path):
            print(f"Config file not found: {path}, using defaults")
            return cfg
        with open(path, "r") as fh:
            data = json.load(fh)
        for section, target in (("ppo", cfg), ("env", cfg.env_config)):
            for key, value in data.get(section, {}).items():
                if hasattr(target, key):
                    setattr(target, key, value)
        fp = data.get("symmetry", {}).get("flip_params", {})
        for block, prefix in (("action_index_info", "flip_action"), ("observation_index_info", "flip_obs")):
            if block in fp:
                setattr(cfg.env_config, prefix + "_right", list(fp[block]["right"]))
                setattr(cfg.env_config, prefix + "_left", list(fp[block]["left"]))
                setattr(cfg.env_config, prefix + "_sign", list(fp[block]["negative_sign"]))
        return cfg
