"""Escape hatch for the unpinned parity: run the REAL reference stack (jax + mujoco + mujoco.mjx, reference
requirements.txt:17-27) on fixed seeds and dump golden vectors that tests/test_golden.py picks up.

Run on any machine that has `pip install mujoco-mjx==3.3.6 jax` and a checkout of son-engr-kr/mujoco-mjx-lab:

    python tools/dump_mjx_golden.py /path/to/mujoco-mjx-lab tests/golden/mjx_humanoid.npz

It cannot run in the build container or on the GPU box (the packages are absent); nothing imports it.
"""
import sys

import numpy as np


def main(ref_root: str, out_path: str, n: int = 64):
    sys.path.insert(0, ref_root)
    import jax
    import jax.numpy as jnp
    import mujoco
    from mujoco import mjx
    from src.config import EnvConfig
    from src.training_utils import load_model_and_create_env

    cfg = EnvConfig(posture_penalty_weight=0.0, random_flip=True)
    m, sys_, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step = load_model_and_create_env(
        f"{ref_root}/models/humanoid_mjx.xml", cfg)
    out = {"nefc": np.array(int(mjx.make_data(sys_).efc_J.shape[0])), "contact_geom": np.asarray(mjx.make_data(sys_).contact.geom)}
    for name in ("body_mass", "body_inertia", "body_ipos", "dof_invweight0", "body_invweight0", "tendon_invweight0", "geom_size",
                 "geom_pos", "geom_quat", "jnt_range", "qpos0"):
        out["model_" + name] = np.asarray(getattr(m, name))
    out["model_meaninertia"] = np.array(m.stat.meaninertia)
    # env layer: reset + 128 steps with fixed actions
    rng = jax.random.PRNGKey(42)
    _, k = jax.random.split(rng)
    keys = jax.random.split(k, n)
    out["keys"] = np.asarray(jax.random.key_data(keys))
    state, obs = v_reset(keys)
    out["reset_qpos"], out["reset_qvel"], out["reset_aux"], out["reset_obs"] = map(np.asarray, (state[0].qpos, state[0].qvel, state[1], obs))
    out["reset_qacc_warmstart"] = np.asarray(state[0].qacc_warmstart)
    acts = np.clip(np.random.default_rng(0).normal(size=(128, n, nu)), -1, 1).astype(np.float32)
    out["actions"] = acts
    traj = {k_: [] for k_ in ("qpos", "qvel", "qacc_warmstart", "time", "aux", "obs", "reward", "terminated", "truncated", "efc_force", "contact_dist")}
    for t in range(128):
        state, obs, r, te, tr = v_step(state, jnp.asarray(acts[t]))
        d, aux = state
        for k_, v in (("qpos", d.qpos), ("qvel", d.qvel), ("qacc_warmstart", d.qacc_warmstart), ("time", d.time), ("aux", aux), ("obs", obs),
                      ("reward", r), ("terminated", te), ("truncated", tr), ("efc_force", d.efc_force), ("contact_dist", d.contact.dist)):
            traj[k_].append(np.asarray(v))
    for k_, v in traj.items():
        out["traj_" + k_] = np.stack(v)
    # speed-test semantics (mjx_humanoid_speed_test.py:48-57)
    vel = jnp.linspace(0.0, 1.0, n)

    def step(v):
        dd = mjx.make_data(sys_)
        dd = dd.replace(qvel=dd.qvel.at[0].set(v))
        return mjx.step(sys_, dd).qpos[0]
    out["speed_vel"], out["speed_pos"] = np.asarray(vel), np.asarray(jax.jit(jax.vmap(step))(vel))
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
