#!/usr/bin/env python
"""North-star CPU baseline B2: MJX on the JAX CPU backend -- BASELINE.json configs[0] to the letter of the reference's
mjx_humanoid_speed_test.py:48-57 (batched step) and :88-103 (device loop), N = 64 envs, humanoid_mjx.xml.

Self-reporting: prints one JSON line.  jax / mujoco-mjx are NOT installed in this image (and there is no network), so here it prints
    {"baseline": "mjx_cpu", "status": "UNAVAILABLE", "why": ...}
and exits 0; on a machine with `pip install jax==0.7.2 mujoco-mjx==3.3.6` it measures.  It uses none of this repository's code.

    python baseline/run_mjx_cpu.py [--xml PATH] [--batch 64] [--iters 10]
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("JAX_PLATFORMS", "cpu")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--xml", default=os.environ.get("HUMANOID_MJX_XML", "models/humanoid_mjx.xml"))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    try:
        import jax
        import jax.numpy as jnp
        import mujoco
        import mujoco.mjx as mjx
    except Exception as e:
        print(json.dumps({"baseline": "mjx_cpu", "status": "UNAVAILABLE", "why": f"module not installed in image ({type(e).__name__}: {e})"}))
        return 0
    if not os.path.exists(args.xml):
        print(json.dumps({"baseline": "mjx_cpu", "status": "UNAVAILABLE", "why": f"{args.xml} not found (pass --xml)"}))
        return 0
    model = mujoco.MjModel.from_xml_path(args.xml)
    mjx_model = mjx.put_model(model)

    def step(vel):                                           # mjx_humanoid_speed_test.py:48-57
        d = mjx.make_data(mjx_model)
        d = d.replace(qvel=d.qvel.at[0].set(vel))
        return mjx.step(mjx_model, d).qpos[0]

    fn = jax.jit(jax.vmap(step))
    vel = jnp.linspace(0.0, 1.0, args.batch)
    fn(vel).block_until_ready()

    @jax.jit
    def repeat_on_device(vel, count):                        # :88-93
        return jax.lax.fori_loop(0, count, lambda i, acc: acc + jnp.sum(fn(vel)), 0.0)

    repeat_on_device(vel, 1).block_until_ready()
    t0 = time.time()
    repeat_on_device(vel, args.iters).block_until_ready()
    dt = max(time.time() - t0, 1e-12)
    print(json.dumps({"baseline": "mjx_cpu", "status": "ok", "backend": jax.default_backend(), "batch": args.batch, "iters": args.iters,
                      "env_steps_per_sec": args.batch * args.iters / dt, "seconds": dt, "cores": os.cpu_count(),
                      "jax": jax.__version__, "mujoco": mujoco.__version__}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
