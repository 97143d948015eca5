#!/usr/bin/env python
"""North-star CPU baseline B1: regular MuJoCo C (`mj_step`) on humanoid.xml, the speed-test workload of the reference's
mjx_humanoid_speed_test.py:48-57 run through the C engine: N independent MjData (make_data; qvel[0] = linspace(0,1,N)[i]; mj_step; read
qpos[0]), `iters` times, on `--threads` host threads (mujoco.rollout-free: one Python thread per shard, the GIL is released in mj_step).

Self-reporting: prints one JSON line.  `mujoco` is NOT installed in this image (and there is no network), so here it prints
    {"baseline": "mujoco_c", "status": "UNAVAILABLE", "why": ...}
and exits 0; on a machine with `pip install mujoco==3.3.6` it measures.  It uses none of this repository's code.

    python baseline/run_mujoco_c.py [--xml PATH] [--batch 64] [--iters 10] [--threads N]
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--xml", default=os.environ.get("HUMANOID_XML", "models/humanoid.xml"))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--threads", type=int, default=len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    args = ap.parse_args()
    try:
        import mujoco
        import numpy as np
    except Exception as e:
        print(json.dumps({"baseline": "mujoco_c", "status": "UNAVAILABLE", "why": f"module not installed in image ({type(e).__name__}: {e})"}))
        return 0
    if not os.path.exists(args.xml):
        print(json.dumps({"baseline": "mujoco_c", "status": "UNAVAILABLE", "why": f"{args.xml} not found (pass --xml)"}))
        return 0
    model = mujoco.MjModel.from_xml_path(args.xml)
    vel = np.linspace(0.0, 1.0, args.batch)
    shards = [list(range(t, args.batch, args.threads)) for t in range(args.threads)]
    datas = [mujoco.MjData(model) for _ in range(args.threads)]

    def run(t):
        d, acc = datas[t], 0.0
        for _ in range(args.iters):
            for i in shards[t]:
                mujoco.mj_resetData(model, d)                # make_data
                d.qvel[0] = vel[i]
                mujoco.mj_step(model, d)
                acc += d.qpos[0]
        return acc

    with ThreadPoolExecutor(args.threads) as ex:
        list(ex.map(run, range(args.threads)))               # warm-up
        t0 = time.time()
        list(ex.map(run, range(args.threads)))
        dt = max(time.time() - t0, 1e-12)
    print(json.dumps({"baseline": "mujoco_c", "status": "ok", "batch": args.batch, "iters": args.iters, "threads": args.threads,
                      "env_steps_per_sec": args.batch * args.iters / dt, "seconds": dt, "mujoco": mujoco.__version__}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
