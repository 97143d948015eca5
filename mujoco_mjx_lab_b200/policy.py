"""Host mirror of the fused policy-inference kernel (csrc/mjxb_policy.cu; include/mjxb.h mjxb_policy_*).

Replaces, inside the rollout loop of reference train_ppo.py:128-169, the chain  normalise -> MLP (src/networks.py:55-61) ->
sample -> log-prob (train_ppo.py:121-126)  by one launch on the tcgen05 tensor cores. Fixed to the reference's default policy shape
(three tanh hidden layers of 256, obs_dim <= 64, act_dim <= 32); anything else raises and the caller keeps its torch path."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib

HID, IN_PAD, OUT_PAD = 256, 64, 32


def supported(params: List[torch.Tensor], obs_dim: int, act_dim: int) -> bool:
    """params: the eight weight / bias tensors; the input layer may carry zero rows beyond obs_dim and the output layer zero columns
    beyond act_dim (the learner's padded copies, ppo.PPOTrainer): the packed image is the same."""
    if len(params) != 8 or obs_dim > IN_PAD or act_dim > OUT_PAD:
        return False
    shapes = [tuple(p.shape) for p in params]
    k0, n3 = shapes[0][0], shapes[6][1] if len(shapes[6]) == 2 else -1
    return (obs_dim <= k0 <= IN_PAD and act_dim <= n3 <= OUT_PAD and all(p.is_contiguous() for p in params) and
            shapes == [(k0, HID), (HID,), (HID, HID), (HID,), (HID, HID), (HID,), (HID, n3), (n3,)])


class FusedPolicy:
    """Packed bf16 image of the policy weights + the launch. `pack()` must be called after every optimiser step (it is four tiny
    kernels on the current stream, capturable in a CUDA graph)."""

    def __init__(self, params: List[torch.Tensor], log_std: torch.Tensor, obs_dim: int, act_dim: int):
        if not supported(params, obs_dim, act_dim):
            raise ValueError("fused policy kernel: unsupported network shape")
        self.params, self.log_std, self.obs_dim, self.act_dim = params, log_std, obs_dim, act_dim
        dev = params[0].device
        self.dims = [(params[0].shape[0], HID, IN_PAD, HID), (HID, HID, HID, HID), (HID, HID, HID, HID), (HID, params[6].shape[1], HID, OUT_PAD)]
        self.packed = [torch.zeros(npad * kpad, dtype=torch.bfloat16, device=dev) for (_, _, kpad, npad) in self.dims]
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        self._w = (C.c_void_p * 4)(*[t.data_ptr() for t in self.packed])
        self._b = (C.c_void_p * 4)(*[params[2 * i + 1].data_ptr() for i in range(4)])
        self.L = _lib.lib()
        self.pack()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def pack(self):
        for i, (k, n, kpad, npad) in enumerate(self.dims):
            w = self.params[2 * i]
            assert w.is_contiguous() and w.dtype == torch.float32
            _lib.check(self.L.mjxb_policy_pack_weight(w.data_ptr(), k, n, kpad, npad, self.packed[i].data_ptr(), self._stream()),
                       "mjxb_policy_pack_weight")

    def act(self, obs: torch.Tensor, eps: torch.Tensor, rms_mean: Optional[torch.Tensor], rms_var: Optional[torch.Tensor],
            act_out: Optional[torch.Tensor] = None, logp_out: Optional[torch.Tensor] = None, mean_out: Optional[torch.Tensor] = None):
        n = obs.shape[0]
        for t in (obs, eps):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError("fused policy kernel: expected contiguous float32 CUDA tensors")
        if obs.shape != (n, self.obs_dim) or eps.shape != (n, self.act_dim):
            raise ValueError("fused policy kernel: shape mismatch")
        act = act_out if act_out is not None else torch.empty(n, self.act_dim, device=obs.device)
        logp = logp_out if logp_out is not None else torch.empty(n, device=obs.device)
        for t in (act, logp) + ((mean_out,) if mean_out is not None else ()):
            if not (t.is_contiguous() and t.dtype == torch.float32):
                raise TypeError("fused policy kernel: outputs must be contiguous float32")
        _lib.check(self.L.mjxb_policy_act(n, self.obs_dim, self.act_dim, obs.data_ptr(),
                                          rms_mean.data_ptr() if rms_mean is not None else None,
                                          rms_var.data_ptr() if rms_var is not None else None,
                                          self._w, self._b, self.log_std.data_ptr(), eps.data_ptr(), act.data_ptr(), logp.data_ptr(),
                                          mean_out.data_ptr() if mean_out is not None else None, self.error.data_ptr(),
                                          self._stream()), "mjxb_policy_act")
        return act, logp
