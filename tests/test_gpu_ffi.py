"""XLA-FFI adapter (csrc/tmjx_xla_ffi.cc): the handler symbols `tmjx_step_ffi` / `tmjx_forward_ffi` are driven through a call frame
built by hand the way XLA's custom-call thunk builds it (`tmjx_ffi_selftest`: metadata probe, a non-execute stage, then EXECUTE with
operands, donated results, int64 attributes and the stream from `XLA_FFI_Stream_Get`), and must give the bits of the direct C-ABI
call.  jaxlib is not installable in this image, so the frame layout comes from tmjx_xla_ffi_c_api_min.h (see its caveat) unless the
real header was found at build time (`tmjx_xla_ffi_available() == 1`)."""
import ctypes as C

import numpy as np
import pytest
import torch

import common
from track_mjx_b200 import _lib as L
from track_mjx_b200 import config
from track_mjx_b200.env import Stepper

pytestmark = pytest.mark.gpu


def _dims(g):
    d = g.dims
    sd = (C.c_int * len(L.STATE_FIELDS))(*[L.field_size(spec, d) for _, spec, _ in L.STATE_FIELDS])
    si = (C.c_int * len(L.STATE_FIELDS))(*[1 if kind == "i" else 0 for _, _, kind in L.STATE_FIELDS])
    od = (C.c_int * 5)(*[L.field_size(spec, d) for _, spec, _ in L.OUT_FIELDS])
    return sd, si, od


def test_handlers_through_a_hand_built_call_frame_match_the_direct_call(walker, clips2, task_cfg):
    lib = L.load()
    assert lib.tmjx_xla_ffi_available() in (0, 1)
    n = 64
    a, b = Stepper(walker.blob, task_cfg, clips2, n, 0), Stepper(walker.blob, task_cfg, clips2, n, 0)
    hb = {k: np.zeros(tuple(v.shape), np.float32 if v.dtype == torch.float32 else np.int32) for k, v in a.buf.items()}
    init = common.init_buffers(hb, clips2, seed=4)
    common.put(a.buf, init); common.put(b.buf, init)
    sd, si, od = _dims(a)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # forward: direct on `a`, through the FFI handler on `b`
    a.forward(L.TMJX_F_SNAPSHOT)
    rc = lib.tmjx_ffi_selftest(0, b._model, b._clips, None, 0, C.byref(b._state_c), C.byref(b._out_c), sd, si, od, n, L.TMJX_F_SNAPSHOT, st, 1)
    assert rc == 0, lib.tmjx_ffi_selftest_error()
    torch.cuda.synchronize()
    for k in ("qpos", "xpos", "obs", "first_obs", "qacc_warmstart", "cur_frame"):
        assert torch.equal(a.buf[k], b.buf[k]), k
    # three steps with the fused wrappers
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(3):
        act = 0.1 * torch.randn(n, walker.nu, device="cuda", generator=g)
        a.step(act, L.TMJX_F_AUTORESET)
        rc = lib.tmjx_ffi_selftest(1, b._model, b._clips, C.c_void_p(act.data_ptr()), walker.nu, C.byref(b._state_c), C.byref(b._out_c), sd, si, od, n,
                                   L.TMJX_F_AUTORESET, st, 1)
        assert rc == 0, lib.tmjx_ffi_selftest_error()
    torch.cuda.synchronize()
    for k in ("qpos", "qvel", "act", "time", "obs", "reward", "done", "metrics", "cur_frame", "action_buffer", "buffer_index", "steps"):
        assert torch.equal(a.buf[k], b.buf[k]), k
    # error paths: state leaves not donated -> FAILED_PRECONDITION (9); a C-ABI error surfaces as INTERNAL (13) with its message
    rc = lib.tmjx_ffi_selftest(1, b._model, b._clips, C.c_void_p(act.data_ptr()), walker.nu, C.byref(b._state_c), C.byref(b._out_c), sd, si, od, n, 0, st, 0)
    assert rc == 9 and b"donated" in lib.tmjx_ffi_selftest_error()
    rc = lib.tmjx_ffi_selftest(1, b._model, None, C.c_void_p(act.data_ptr()), walker.nu, C.byref(b._state_c), C.byref(b._out_c), sd, si, od, n, 0, st, 1)
    assert rc == 13 and b"null" in lib.tmjx_ffi_selftest_error()
    a.close(); b.close()
