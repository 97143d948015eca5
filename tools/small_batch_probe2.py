import sys, os, json, numpy as np, torch
sys.path[:0] = ['.', 'tests']
import helpers
from mujoco_mjx_lab_b200 import _lib, training_utils, parallel
import bench
model = helpers.load()
out = {}
for name, flags in (("spec", None), ("deferred", _lib.FLAG_NO_SPEC_RESET)):
    env = training_utils.load_model_and_create_env("", helpers.env_config(), model=model, flags=flags)
    out[name] = [bench.sweep_point(n, env[8], env[9], env[9].sys, 21, torch.device("cuda:0")) for n in (64, 512, 1024, 1184, 2048, 4096)]
    for r in out[name]:
        print(name, r["n_env"], "graph us/step %.1f" % (r["trajectory_graph_ms_per_step"] * 1e3), "eager %.1f" % (r["trajectory_eager_ms_per_step"] * 1e3), "speedtest %.1f" % (r["speedtest_ms_per_step"] * 1e3), flush=True)
json.dump(out, open("gpurun_out/r2_small_batch.json", "w"), indent=1)
