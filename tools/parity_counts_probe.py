import sys, numpy as np, torch
sys.path[:0] = ['.', 'tests']
import helpers
from mujoco_mjx_lab_b200 import mjx, training_utils
from oracle import oracle as O
O.build()
model = helpers.load(); cfg = helpers.env_config()
orc = helpers.make_oracle(model, cfg)
env = training_utils.load_model_and_create_env("", cfg, model=model)
N = lambda t: t.detach().double().cpu().numpy()
T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")
for n, seed in ((1024, 42), (4096, 7)):
    keys = helpers.ppo_keys(seed, n)
    (d, aux), obs = env[8](keys)
    st, o_ref = orc.env_reset(keys, prec="f32")
    s64, _ = orc.env_reset(keys, prec="f64")
    print("reset n", n, "stance mismatch vs o32", int((N(aux)[:, 5] != st["aux"][:, 5]).sum()), "o32 vs o64", int((st["aux"][:, 5] != s64["aux"][:, 5]).sum()),
          "gpu vs o64", int((N(aux)[:, 5] != s64["aux"][:, 5]).sum()), "max |obs - o32|", float(np.abs(N(obs) - o_ref).max()),
          "max|warm - o64| gpu", float(np.abs(N(d.qacc_warmstart) - s64["qacc_warmstart"]).max()), "o32", float(np.abs(st["qacc_warmstart"] - s64["qacc_warmstart"]).max()))
sysm = mjx.put_model(model)
for kind in ["free", "stand", "lean", "tumble"]:
    q, v, w, c = helpers.make_states(model, 512, 100 + ["free", "stand", "lean", "tumble"].index(kind), kind)
    ref = orc.forward(q, v, w, c, prec="f64", debug=True)
    r32 = orc.forward(q, v, w, c, prec="f32", debug=True)
    _, out = mjx.forward(sysm, mjx.Data(T(q), T(v), T(w), torch.zeros(512, device="cuda"), T(c)), debug=True)
    g = {k: (t.cpu().numpy() if t.dtype == torch.int32 else N(t)) for k, t in out.items()}
    cm = (g["efc_active"] & 1) != (ref["efc_active"] & 1)
    am = ((g["efc_active"] >> 1) != (ref["efc_active"] >> 1)) & ~cm
    am32 = ((r32["efc_active"] >> 1) != (ref["efc_active"] >> 1))
    sg = (g["sensordata"] > 0) != (ref["sensordata"] > 0)
    s32 = (r32["sensordata"] > 0) != (ref["sensordata"] > 0)
    nd = mjx.step(sysm, mjx.Data(T(q), T(v), T(w), torch.zeros(512, device="cuda"), T(c)))
    r64s = orc.physics_step(q, v, w, None, c, prec="f64"); r32s = orc.physics_step(q, v, w, None, c, prec="f32")
    ew = np.abs(N(nd.qacc_warmstart) - r64s["qacc_warmstart"]).max(axis=1); ew32 = np.abs(r32s["qacc_warmstart"] - r64s["qacc_warmstart"]).max(axis=1)
    ev = np.abs(N(nd.qvel) - r64s["qvel"]).max(axis=1); ev32 = np.abs(r32s["qvel"] - r64s["qvel"]).max(axis=1)
    print(kind, "cand mism vs o64", int(cm.sum()), "active mism gpu", int(am.sum()), "o32", int(am32.sum()), "touch sign mism gpu", int(sg.sum()), "o32", int(s32.sum()),
          "| qacc_ws median gpu %.2e o32 %.2e max gpu %.2e o32 %.2e" % (np.median(ew), np.median(ew32), ew.max(), ew32.max()),
          "| qvel median gpu %.2e o32 %.2e max gpu %.2e o32 %.2e" % (np.median(ev), np.median(ev32), ev.max(), ev32.max()))
