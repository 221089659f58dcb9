"""Python driver of the CPU oracle (TEST INFRASTRUCTURE — see the header of tmjx_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Buffers are numpy arrays with the same names and `[n_env, dim]` layout as the device buffers of the
product path, so a parity test feeds both sides from one dict.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from track_mjx_b200 import _lib as L  # noqa: E402  (struct layouts only; libtmjx.so is NOT loaded here)
from track_mjx_b200.config import TaskConfigC, TMJX_N_METRICS  # noqa: E402

LIB_PATH = os.path.join(_HERE, "libtmjx_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tmjx_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "libtmjx_oracle.so"], check=True,
                       capture_output=True)
    return LIB_PATH


class Oracle:
    """fp32 / fp64 CPU restatement of the reference env step for one (model, task config, clip table)."""

    def __init__(self, blob: bytes, cfg: TaskConfigC, clips, *, dtype=np.float32, nthreads: int | None = None):
        build()
        self.lib = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        fp = C.POINTER(C.c_float)
        self.lib.tmjx_oracle_last_error.restype = C.c_char_p
        self.lib.tmjx_oracle_create.argtypes = [vp, C.c_size_t, C.POINTER(TaskConfigC), C.POINTER(vp)]
        self.lib.tmjx_oracle_destroy.argtypes = [vp]
        self.lib.tmjx_oracle_set_clips.argtypes = [vp] + [fp] * 5 + [C.c_int] * 3
        self.lib.tmjx_oracle_obs_size.argtypes = [vp]
        self.lib.tmjx_oracle_forward.argtypes = [vp, C.POINTER(L.StateC), C.POINTER(L.OutC), C.c_int, C.c_uint, C.c_int, C.c_int]
        self.lib.tmjx_oracle_step.argtypes = [vp, vp, C.POINTER(L.StateC), C.POINTER(L.OutC), C.c_int, C.c_uint, C.c_int, C.c_int]
        self.h = vp()
        self.cfg = cfg
        rc = self.lib.tmjx_oracle_create(blob, len(blob), C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(self.lib.tmjx_oracle_last_error().decode())
        self.dtype = np.dtype(dtype)
        self.nthreads = nthreads or os.cpu_count() or 1
        self._clips = [np.ascontiguousarray(getattr(clips, k), np.float32)
                       for k in ("position", "quaternion", "joints", "body_positions", "angular_velocity")]
        n_clips, clip_len = clips.position.shape[:2]
        self.lib.tmjx_oracle_set_clips(self.h, *[a.ctypes.data_as(fp) for a in self._clips], n_clips, clip_len,
                                       clips.body_positions.shape[2])
        from track_mjx_b200 import model_blob
        d = model_blob.unpack(blob)["dims"]
        self.dims = dict(nq=int(d[0]), nv=int(d[1]), nu=int(d[2]), na=int(d[3]), nbody=int(d[4]), njnt=int(d[5]),
                         ncon=int(d[7]), nefc=int(d[8]), obs_size=int(self.lib.tmjx_oracle_obs_size(self.h)),
                         var_window_size=int(cfg.var_window_size), n_metrics=TMJX_N_METRICS)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.tmjx_oracle_destroy(self.h)
            self.h = None

    def alloc(self, n_env: int, debug: bool = True) -> dict:
        buf = {}
        fields = L.STATE_FIELDS + L.OUT_FIELDS + (L.DEBUG_FIELDS if debug else ())
        for name, spec, kind in fields:
            buf[name] = np.zeros((n_env, L.field_size(spec, self.dims)), self.dtype if kind == "f" else np.int32)
        return buf

    def _structs(self, buf):
        ptr = lambda a: a.ctypes.data  # noqa: E731
        s = L.fill_struct(L.StateC(), L.STATE_FIELDS, buf, ptr)
        o = L.fill_struct(L.OutC(), L.OUT_FIELDS + L.DEBUG_FIELDS, buf, ptr)
        return s, o

    def _code(self):
        return 0 if self.dtype == np.float32 else 1

    def forward(self, buf: dict, flags: int = 0):
        s, o = self._structs(buf)
        n = buf["qpos"].shape[0]
        rc = self.lib.tmjx_oracle_forward(self.h, C.byref(s), C.byref(o), n, flags, self._code(), self.nthreads)
        if rc != 0:
            raise RuntimeError(self.lib.tmjx_oracle_last_error().decode())

    def step(self, buf: dict, action: np.ndarray, flags: int = 0):
        s, o = self._structs(buf)
        n = buf["qpos"].shape[0]
        a = np.ascontiguousarray(action, self.dtype)
        rc = self.lib.tmjx_oracle_step(self.h, a.ctypes.data, C.byref(s), C.byref(o), n, flags, self._code(), self.nthreads)
        if rc != 0:
            raise RuntimeError(self.lib.tmjx_oracle_last_error().decode())
