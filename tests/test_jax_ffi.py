"""The XLA-FFI adapter (csrc/mjxb_ffi.cc + mujoco_mjx_lab_b200/jax_ffi.py) that binds the C ABI into the reference's JAX code
(reference src/envs.py:494-497, train_ppo.py:166-168, train_apg.py:161-209). jax is absent from this image, so the GPU leg skips;
what can be checked without jax is checked: every C-ABI call the handlers make exists with that arity in include/mjxb.h."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ffi_source_calls_only_declared_abi_functions():
    src = open(os.path.join(ROOT, "mujoco_mjx_lab_b200", "csrc", "mjxb_ffi.cc")).read()
    hdr = open(os.path.join(ROOT, "include", "mjxb.h")).read()
    called = set(re.findall(r"\b(mjxb_[a-z_]+)\(", src)) - {"mjxb_ffi"}
    assert {"mjxb_reset", "mjxb_step", "mjxb_step_autoreset", "mjxb_step_vjp"} <= called
    for fn in called:
        assert re.search(r"\b%s\(" % fn, hdr), f"{fn} is not declared in include/mjxb.h"
    # handler symbols the Python side registers
    py = open(os.path.join(ROOT, "mujoco_mjx_lab_b200", "jax_ffi.py")).read()
    for sym in re.findall(r'"(Mjxb[A-Za-z]+)"', py):
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({sym}," in src, sym


@pytest.mark.gpu
def test_jax_env_matches_torch_env():
    jax = pytest.importorskip("jax", reason="jax / jaxlib are not installed in this image (no network): the XLA-FFI adapter ships untested")
    import numpy as np
    import helpers
    from mujoco_mjx_lab_b200 import jax_ffi, training_utils
    model, cfg = helpers.load(), helpers.env_config()
    _, _, v_reset_j, v_step_j = jax_ffi.create_env_functions(model, cfg, model["qpos0"], 28, 27)
    env = training_utils.load_model_and_create_env("", cfg, model=model)
    keys = helpers.ppo_keys(3, 64)
    (dj, auxj), obsj = v_reset_j(jax.numpy.asarray(keys))
    (dt, auxt), obst = env[8](keys)
    np.testing.assert_array_equal(np.asarray(obsj), obst.cpu().numpy())
    act = np.random.default_rng(0).normal(size=(64, 21)).astype(np.float32)
    import torch
    outj = jax.jit(v_step_j)((dj, auxj), jax.numpy.asarray(act))
    outt = env[9]((dt, auxt), torch.from_numpy(act).cuda())
    np.testing.assert_array_equal(np.asarray(outj[1]), outt[1].cpu().numpy())
    np.testing.assert_array_equal(np.asarray(outj[2]), outt[2].cpu().numpy())
