"""threefry2x32 known-answer tests (Random123 / JAX random_test vectors) for both restatements."""
import numpy as np

from mujoco_mjx_lab_b200 import jax_random
from oracle import oracle as O

KATS = [((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6b200159, 0x99ba4efe)),
        ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
        ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0))]


def test_threefry_kats_oracle():
    for key, ctr, out in KATS:
        assert O.threefry2x32(key[0], key[1], ctr[0], ctr[1]) == out


def test_threefry_kats_host():
    for key, ctr, out in KATS:
        x0, x1 = jax_random.threefry2x32(np.array(key, dtype=np.uint32), np.array([ctr[0]], dtype=np.uint32), np.array([ctr[1]], dtype=np.uint32))
        assert (int(x0[0]), int(x1[0])) == out


def test_split_uniform_agree_between_host_and_oracle():
    key = jax_random.PRNGKey(42)
    np.testing.assert_array_equal(jax_random.split(key, 7), O.jax_split(key, 7))
    np.testing.assert_array_equal(jax_random.uniform(key, 33, -1.0, 2.5), O.jax_uniform(key, 33, -1.0, 2.5))
    u = jax_random.uniform(key, 4096)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.03
    assert list(jax_random.PRNGKey(42)) == [0, 42]
