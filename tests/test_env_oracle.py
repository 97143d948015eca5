"""Env layer (reference src/envs.py) restated in the oracle: structure, RNG use, flip maps, done logic."""
import numpy as np

import helpers
from mujoco_mjx_lab_b200 import _abi, config, jax_random


def test_flip_permutations_match_reference_construction():
    cfg = helpers.env_config()
    ap, asg, op, osg = _abi.flip_permutations(cfg, 21, 54)
    # reference src/envs.py:61-74 with the index lists of src/config.py:60-66
    r, l = [3, 4, 5, 6, 7, 8, 15, 16, 17], [9, 10, 11, 12, 13, 14, 18, 19, 20]
    exp = np.arange(21); exp[r], exp[l] = l, r
    np.testing.assert_array_equal(ap, exp)
    assert list(np.nonzero(asg < 0)[0]) == [0, 2]
    assert list(np.nonzero(osg < 0)[0]) == [1, 3, 4, 6, 26, 28, 30, 31, 33, 52]
    assert op[7] == 13 and op[13] == 7 and op[46] == 49 and op[51] == 48 and op[0] == 0
    # involution: applying the flip twice is the identity
    x = np.random.default_rng(0).normal(size=54)
    np.testing.assert_allclose((x[op] * osg)[op] * osg, x)
    cfg2 = helpers.env_config(random_flip=False)
    ap2, asg2, op2, osg2 = _abi.flip_permutations(cfg2, 21, 54)
    assert (ap2 == np.arange(21)).all() and (osg2 == 1).all()


def test_out_of_range_flip_indices_rejected():
    cfg = helpers.env_config()
    cfg.flip_obs_right = [29, 79]      # reference src/config_test.json:56-93 style indices
    cfg.flip_obs_left = [30, 78]
    try:
        _abi.flip_permutations(cfg, 21, 54)
    except ValueError:
        return
    raise AssertionError("expected ValueError")


def test_ppo_config_json_overlay(tmp_path):
    p = tmp_path / "c.json"
    p.write_text('{"env": {"posture_penalty_weight": 0.0, "random_flip": true, "id": "x"}, "ppo": {"num_envs": 1024, "bogus": 1},'
                 ' "symmetry": {"flip_params": {"action_index_info": {"right": [3], "left": [9], "negative_sign": [0]}}}}')
    c = config.PPOConfig.from_json(str(p))
    assert c.num_envs == 1024 and c.env_config.random_flip is True and c.env_config.posture_penalty_weight == 0.0
    assert c.env_config.flip_action_right == [3] and not hasattr(c, "bogus")
    assert config.PPOConfig.from_json(str(tmp_path / "missing.json")).rollout_length == 128


def test_reset_semantics(model, oracle):
    keys = helpers.ppo_keys(42, 64)
    st, obs = oracle.env_reset(keys, prec="f32")
    st2, obs2 = oracle.env_reset(keys, prec="f32")
    np.testing.assert_array_equal(obs, obs2)
    # noise amplitudes (src/envs.py:127-131) and the uniform draws themselves
    for e in (0, 17):
        k1, k2, k3, k4 = jax_random.split(keys[e], 4)
        nz = jax_random.uniform(k1, 21) * np.float32(2) - np.float32(1)
        np.testing.assert_array_equal(st["qpos"][e, 7:].astype(np.float32), (np.float32(0) + np.float32(0.01) * nz))
        nv_ = jax_random.uniform(k2, 27) * np.float32(2) - np.float32(1)
        np.testing.assert_array_equal(st["qvel"][e, 2:].astype(np.float32), (np.float32(0.01) * nv_)[2:])
        assert st["aux"][e, 0] == float(jax_random.uniform(k3, 1)[0] < 0.5)
        vmag = jax_random.uniform(k4, 1, 0.0, 0.5)[0]
        assert abs(st["qvel"][e, 0] - vmag) < 1e-7 and st["qvel"][e, 1] == 0.0
    assert set(np.unique(st["aux"][:, 0])) == {0.0, 1.0}
    np.testing.assert_allclose(st["qpos"][:, :7], np.tile(model["qpos0"][:7], (64, 1)))
    aux = st["aux"]
    np.testing.assert_allclose(aux[:, 1], -0.01 + 2.0, atol=5e-3)       # tx = pelvis_x + target_dist (pelvis moves with the abdomen noise)
    np.testing.assert_allclose(aux[:, 3], 0.857, atol=2e-3)            # tz = pelvis height
    assert (aux[:, 4] == 0).all() and (aux[:, 8] == 0).all() and (st["time"] == 0).all()
    assert np.isin(aux[:, 5], [0, 1, 2, 3]).all()
    np.testing.assert_allclose(aux[:, 7], -2.0 / 0.005, rtol=5e-3)
    # obs layout (src/envs.py:193-200): height, rpy, joint pos, local vel, target feature
    unflipped = aux[:, 0] == 0
    np.testing.assert_allclose(obs[unflipped, 0], 0.857, atol=2e-3)
    np.testing.assert_allclose(obs[unflipped][:, 4:25], st["qpos"][unflipped][:, 7:], atol=1e-7)
    np.testing.assert_allclose(obs[unflipped][:, 31:52], st["qvel"][unflipped][:, 6:], atol=1e-7)
    # warm start is the forward pass's solution (mjx.forward writes qacc_warmstart)
    assert np.abs(st["qacc_warmstart"]).max() > 1.0


def test_step_semantics(model, oracle):
    keys = helpers.ppo_keys(7, 32)
    st, obs = oracle.env_reset(keys, prec="f32")
    rng = np.random.default_rng(0)
    act = rng.normal(size=(32, 21))
    st1, obs1, r, te, tr, mask, _ = oracle.env_step(st, act, prec="f32")
    np.testing.assert_allclose(st1["time"], 0.005, rtol=1e-6)
    assert (st1["aux"][:, 8] == 1).all() and (te == 0).all() and (tr == 0).all() and (mask == 0).all()
    assert (st1["aux"][:, 0] == st["aux"][:, 0]).all()
    # progress term dominates: reward ~ (pot' - pot) - energy; potential is -dist/dt
    assert np.isfinite(r).all() and np.abs(r).max() < 50
    # flipped envs see the mirrored action: stepping with the pre-mirrored action equals an unflipped env's step
    ap, asg, op, osg = _abi.flip_permutations(helpers.env_config(), 21, 54)
    st_nf = {k: v.copy() for k, v in st.items()}
    st_nf["aux"][:, 0] = 0.0
    act_m = np.where(st["aux"][:, :1] > 0.5, act[:, ap] * asg, act)
    st2, obs2, r2, *_ = oracle.env_step(st_nf, act_m, prec="f32")
    np.testing.assert_array_equal(st2["qpos"], st1["qpos"])
    np.testing.assert_array_equal(r2, r)
    flipped = st["aux"][:, 0] > 0.5
    np.testing.assert_array_equal(obs1[flipped], (obs2[flipped][:, op] * osg))
    np.testing.assert_array_equal(obs1[~flipped], obs2[~flipped])


def test_truncation_termination_and_autoreset(model, oracle):
    keys = helpers.ppo_keys(3, 8)
    st, _ = oracle.env_reset(keys, prec="f32")
    st["aux"][:4, 8] = 998.0                       # next step reaches max_episode_steps = 1000? no: 999
    st["aux"][4:, 8] = 999.0                       # -> 1000 => truncated
    st["qpos"][0, 2] = 0.9                         # pelvis at 0.475 < 0.7 => terminated
    act = np.zeros((8, 21))
    rk = helpers.ppo_keys(99, 8)
    st1, obs1, r, te, tr, mask, _ = oracle.env_step(st, act, prec="f32", reset_keys=rk)
    assert list(tr) == [0, 0, 0, 0, 1, 1, 1, 1] and list(te) == [1, 0, 0, 0, 0, 0, 0, 0]
    assert list(mask) == [1, 0, 0, 0, 1, 1, 1, 1]
    fresh, fobs = oracle.env_reset(rk, prec="f32")
    done = mask.astype(bool)
    for k in st1:
        np.testing.assert_array_equal(st1[k][done], fresh[k][done])
    np.testing.assert_array_equal(obs1[done], fobs[done])
    assert (st1["aux"][~done, 8] == 999).all()
