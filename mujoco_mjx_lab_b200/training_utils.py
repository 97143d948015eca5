"""Mirror of the reference loader (reference src/training_utils.py:70-114): the one place where the backend is chosen."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from . import mjx, modelc
from .envs import create_env_functions


def load_model_and_create_env(xml_path: str, env_config: Any, lighten_solver: bool = False,
                              solver_options: Optional[Dict[str, Any]] = None, model: Optional[Dict[str, Any]] = None,
                              variant: str = "fast", flags: Optional[int] = None):
    """Returns (m, sys, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step) like the reference.

    `m` is the compiled-constants dict (stands in for mujoco.MjModel), `sys` the device model (stands in for mjx.Model).
    `model` may pass an already compiled dict (e.g. data/humanoid_mjx.json) instead of an XML path.
    """
    overrides: Dict[str, Any] = {}
    if lighten_solver:  # reference :95-98
        overrides.update(iterations=1, ls_iterations=1)
    if solver_options:  # reference :100-103
        overrides.update(solver_options)
    if model is None:
        print(f"Loading model: {xml_path}")
        m = modelc.compile_mjcf(xml_path, overrides or None)
    else:
        m = dict(model)
        m["opt"] = dict(m["opt"], **overrides)
    env_config.pelvis_body_id = m["body_name"].index("pelvis")
    env_config.head_body_id = m["body_name"].index("head")
    env_config.touch_sensor_right_id = m["sensor_name"].index("touch_foot_right")
    env_config.touch_sensor_left_id = m["sensor_name"].index("touch_foot_left")
    sys = mjx.put_model(m, variant=variant, flags=flags)   # variant / flags: test-only knobs (mjx.Model)
    nq, nv, nu = int(m["nq"]), int(m["nv"]), int(m["nu"])
    q0 = np.asarray(m["qpos0"], dtype=np.float32).copy()
    single_reset, single_step, v_reset, v_step = create_env_functions(sys, env_config, q0, nq, nv)
    return m, sys, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step
