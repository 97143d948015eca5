"""Measures the parity table (tests/parity_measure.py) on the GPU and writes profiles/r2_parity.json.

    python tools/parity_table.py [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import helpers  # noqa: E402
import parity_measure as PM  # noqa: E402
from mujoco_mjx_lab_b200 import _lib, training_utils  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_parity.json")
    O.build()
    model = helpers.load()
    cfg = helpers.env_config()
    oracle = helpers.make_oracle(model, cfg)
    syss = PM.systems(model)
    table = {"note": "one physics step from identical states, 512 envs per state family; tolerances are SURVEY.md section 7's "
                     "(see tests/parity_measure.py); *_over_tol = error / tolerance; counts are elements (n_exceed) and envs (n_env_exceed)",
             "one_step": {}}
    for kind in PM.KINDS:
        table["one_step"][kind] = PM.one_step_table(model, oracle, syss, kind)
        e = table["one_step"][kind]["exact_vs_o32"]
        print(kind, "exact vs o32: qvel max/tol %.3g qacc max/tol %.3g efc_force max/tol %.3g niter equal %d/%d active mism %d (away %s)" % (
            e["qvel"]["max_over_tol"], e["qacc"]["max_over_tol"], e["efc_force"]["max_over_tol"], e["solver_niter"]["equal"], e["solver_niter"]["n"],
            e["active_mask"]["mismatch"], e["active_mask"].get("mismatch_away")), flush=True)
    envs = {}
    for name, kw in (("fast", {}), ("fast_it", dict(flags=_lib.FLAG_LS_ITERATIVE)), ("exact", dict(variant="exact"))):
        env = training_utils.load_model_and_create_env("", helpers.env_config(), model=model, **kw)
        envs[name] = (env[9], env[9].sys)
    table["resynchronised_128_env_steps_vs_o32"] = PM.resync_table(model, oracle, envs)
    with open(out_path, "w") as f:
        json.dump(table, f, indent=1)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
