"""Learner-side pieces that sit next to the acting loop (only GAE is built so far).

`compute_gae` mirrors `track_mjx/agent/mlp_ppo/losses.py:39-101`: same argument names and meaning, time-major `[T, B]` fp32 CUDA
tensors in, `(vs, advantages)` out, computed by the `tmjx_gae` kernel (csrc/tmjx_policy.cu).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L


def compute_gae(truncation, termination, rewards, values, bootstrap_value, lambda_: float = 1.0, discount: float = 0.99):
    import torch

    if not rewards.is_cuda:
        raise RuntimeError("compute_gae needs CUDA tensors: there is no CPU fallback")
    T, B = rewards.shape
    args = [a.to(torch.float32).contiguous() for a in (truncation, termination, rewards, values, bootstrap_value)]
    for a in args[:4]:
        if a.shape != (T, B):
            raise ValueError("truncation, termination, rewards and values must all be [T, B]")
    if args[4].shape != (B,):
        raise ValueError("bootstrap_value must be [B]")
    vs, adv = torch.empty_like(args[2]), torch.empty_like(args[2])
    lib = L.load()
    ptr = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.tmjx_gae(*[ptr(a) for a in args], float(lambda_), float(discount), ptr(vs), ptr(adv), int(T), int(B),
                      C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"tmjx_gae failed ({rc}): {lib.tmjx_policy_last_error().decode()}")
    return vs, adv
