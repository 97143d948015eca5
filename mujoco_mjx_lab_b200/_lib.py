"""ctypes loader of the C-ABI shared library (mujoco_mjx_lab_b200/libmjxb.so, built from csrc/ for sm_100a).

There is no CPU fallback: a missing library or a missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from ._abi import DebugC, EnvConfigC, StateC

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MJXB_LIB", os.path.join(_HERE, "libmjxb.so"))  # MJXB_LIB: alternative build (tuning experiments)
# the reference-arithmetic build of the same sources (no fast-math, no FMA contraction, MJX's iterative line search): the yardstick the
# parity tests measure the product against (csrc/Makefile); never a fallback
EXACT_LIB_PATH = os.path.join(_HERE, "libmjxb_exact.so")
FLAG_LS_ITERATIVE, FLAG_DENSE_CHOL, FLAG_INLINE_RESET, FLAG_NO_SPEC_RESET, FLAG_NO_WORK_SORT, FLAG_NO_DYN_ROUNDS, FLAG_BUILD_EXACT = 1, 2, 4, 8, 16, 32, 256

ERRORS = {0: "ok", -1: "invalid argument", -2: "bad model blob", -3: "CUDA error", -4: "no CUDA device (no CPU fallback)",
          -5: "unsupported model"}

# every symbol include/mjxb.h declares (tests check that the built library exports all of them)
SYMBOLS = ("mjxb_abi_version", "mjxb_launch_count", "mjxb_blob_sizeof", "mjxb_env_config_sizeof", "mjxb_strerror", "mjxb_last_cuda_error",
           "mjxb_model_create", "mjxb_model_create_ex", "mjxb_model_flags", "mjxb_model_reserve", "mjxb_ffma_peak", "mjxb_model_destroy", "mjxb_model_dims", "mjxb_model_scratch_bytes", "mjxb_launch_config", "mjxb_reset", "mjxb_step",
           "mjxb_step_autoreset", "mjxb_physics_step", "mjxb_forward", "mjxb_speed_test", "mjxb_reset_host", "mjxb_step_host",
           "mjxb_step_autoreset_host", "mjxb_state_get_host", "mjxb_state_set_host", "mjxb_policy_pack_weight", "mjxb_policy_act", "mjxb_gae", "mjxb_tanh_bwd_colsum",
           "mjxb_step_fwd_tape", "mjxb_step_vjp", "mjxb_ppo_loss", "mjxb_ppo_loss_ld", "mjxb_adam",
           "mjxb_comm_create", "mjxb_comm_local_handles", "mjxb_comm_connect", "mjxb_comm_grad_buffer", "mjxb_comm_error",
           "mjxb_allreduce_adam", "mjxb_comm_destroy")


class MjxbError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ -> libmjxb.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    inc = os.path.join(_HERE, "..", "include")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(inc, f) for f in ("mjxb.h", "mjxb_model.h")]
    for target in (LIB_PATH, EXACT_LIB_PATH):
        stale = force or not os.path.exists(target) or any(os.path.getmtime(s) > os.path.getmtime(target) for s in srcs)
        if stale:
            r = subprocess.run(["make", "-C", csrc, "-j8"] + (["-B"] if force else []) + ["../" + os.path.basename(target)], capture_output=True, text=True)
            if verbose or r.returncode != 0:
                print(r.stdout[-4000:], r.stderr[-4000:])
            if r.returncode != 0:
                raise MjxbError("nvcc build of %s failed" % os.path.basename(target))
    return LIB_PATH


_libs = {}


def lib(variant: str = "fast") -> C.CDLL:
    """The C-ABI library: "fast" = the product (libmjxb.so), "exact" = the reference-arithmetic build (libmjxb_exact.so)."""
    if variant in _libs:
        return _libs[variant]
    path = {"fast": LIB_PATH, "exact": EXACT_LIB_PATH}[variant]
    if not os.path.exists(path):
        raise MjxbError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA extension is mandatory; there is no CPU fallback)")
    L = C.CDLL(path)
    L.mjxb_abi_version.restype = C.c_int
    L.mjxb_launch_count.restype = C.c_longlong
    L.mjxb_blob_sizeof.restype = C.c_size_t
    L.mjxb_env_config_sizeof.restype = C.c_size_t
    L.mjxb_model_scratch_bytes.restype = C.c_size_t
    L.mjxb_model_scratch_bytes.argtypes = [C.c_void_p]
    L.mjxb_strerror.restype = C.c_char_p
    L.mjxb_strerror.argtypes = [C.c_int]
    L.mjxb_last_cuda_error.restype = C.c_char_p
    L.mjxb_model_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(EnvConfigC), C.c_int, C.POINTER(C.c_void_p)]
    L.mjxb_model_create_ex.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(EnvConfigC), C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]
    L.mjxb_model_flags.argtypes = [C.c_void_p]
    L.mjxb_model_reserve.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    L.mjxb_ffma_peak.argtypes = [C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.mjxb_model_destroy.argtypes = [C.c_void_p]
    L.mjxb_model_destroy.restype = None
    L.mjxb_model_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.mjxb_launch_config.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    vp, i32 = C.c_void_p, C.c_int32
    L.mjxb_reset.argtypes = [vp, i32, vp, StateC, vp, vp, vp]
    L.mjxb_step.argtypes = [vp, i32, StateC, vp, StateC, vp, vp, vp, vp, vp, vp]
    L.mjxb_step_autoreset.argtypes = [vp, i32, StateC, vp, vp, StateC, vp, vp, vp, vp, vp, vp, vp]
    L.mjxb_step_fwd_tape.argtypes = [vp, i32, StateC, vp, StateC, vp, vp, vp, vp, vp, vp, vp]
    L.mjxb_step_vjp.argtypes = [vp, i32, StateC, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.mjxb_physics_step.argtypes = [vp, i32, StateC, vp, i32, C.POINTER(DebugC), vp, vp]
    L.mjxb_forward.argtypes = [vp, i32, StateC, vp, C.POINTER(DebugC), vp, vp]
    L.mjxb_speed_test.argtypes = [vp, i32, vp, vp, i32, vp]
    L.mjxb_reset_host.argtypes = [vp, i32, vp, vp]
    L.mjxb_step_host.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.mjxb_step_autoreset_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.mjxb_state_get_host.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.mjxb_state_set_host.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.mjxb_policy_pack_weight.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.mjxb_tanh_bwd_colsum.argtypes = [i32, i32, vp, vp, vp, vp, vp]
    f32 = C.c_float
    L.mjxb_ppo_loss.argtypes = [i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp]
    L.mjxb_ppo_loss_ld.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp]
    L.mjxb_adam.argtypes = [i32, i32, vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, f32, vp]
    L.mjxb_comm_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
    L.mjxb_comm_local_handles.argtypes = [vp, vp]
    L.mjxb_comm_connect.argtypes = [vp, vp]
    L.mjxb_comm_grad_buffer.argtypes = [vp]
    L.mjxb_comm_grad_buffer.restype = vp
    L.mjxb_comm_error.argtypes = [vp]
    L.mjxb_allreduce_adam.argtypes = [vp, i32, i32, vp, vp, vp, vp, f32, f32, f32, f32, f32, vp]
    L.mjxb_comm_destroy.argtypes = [vp]
    L.mjxb_comm_destroy.restype = None
    L.mjxb_gae.argtypes = [i32, i32, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, vp]
    L.mjxb_policy_act.argtypes = [i32, i32, i32, vp, vp, vp, C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp, vp, vp, vp]
    if L.mjxb_abi_version() != 1:
        raise MjxbError("libmjxb.so ABI version mismatch")
    _libs[variant] = L
    return L


def check(rc: int, what: str = "mjxb call", L=None):
    if rc != 0:
        L = L or lib()
        msg = L.mjxb_strerror(rc).decode()
        if rc == -3:
            msg += ": " + L.mjxb_last_cuda_error().decode()
        raise MjxbError(f"{what} failed: {msg} ({rc})")
