"""CPU-side checks: the C-ABI library loads and exports every symbol include/tmjx.h declares, struct layouts
agree between ctypes and C, blob round-trips, index tables, config arithmetic (no compute calls: no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from track_mjx_b200 import _lib as L
from track_mjx_b200 import clips as clipmod, config, model_blob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    header = open(os.path.join(ROOT, "include", "tmjx.h")).read()
    declared = set(re.findall(r"\b(tmjx_[a-z0-9_]+)\s*\(", header))
    assert {"tmjx_step", "tmjx_forward", "tmjx_model_create", "tmjx_clips_create"} <= declared
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tmjx_abi_version() == config.TMJX_ABI_VERSION


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "tmjx.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(TmjxTaskConfig), '
                   'sizeof(TmjxState), sizeof(TmjxOut), sizeof(TmjxDims), __builtin_offsetof(TmjxTaskConfig, torso_idx));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(config.TaskConfigC), C.sizeof(L.StateC), C.sizeof(L.OutC), C.sizeof(L.DimsC),
                     config.TaskConfigC.torso_idx.offset]


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libtmjx.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        L.load()


def test_blob_roundtrip_and_dims(walker):
    sec = model_blob.unpack(walker.blob)
    again = model_blob.unpack(model_blob.pack_sections(sec))
    assert sec.keys() == again.keys()
    for k in sec:
        assert (sec[k] == again[k]).all()
    assert (walker.nq, walker.nv, walker.nu, walker.na, walker.nbody, walker.njnt) == (74, 73, 38, 38, 68, 68)
    assert (walker.ncon, walker.nefc) == (30, 187)
    assert abs(float(sec["opt"][0]) - 0.002) < 1e-9


def test_walker_index_tables(walker):
    """SURVEY A.4 (mj_name2id equivalents for rodent-full-clips.yaml's name lists)."""
    assert list(walker.joint_idxs) == [1, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 33, 47, 48, 50, 51, 52, 53, 54, 55, 56, 57,
                                       58, 60, 61, 62, 63, 64, 65, 66, 67]
    assert list(walker.body_idxs) == [3, 10, 11, 12, 13, 15, 16, 17, 56, 57, 58, 59, 60, 62, 63, 64, 65, 67]
    assert list(walker.endeff_idxs) == [13, 17, 61, 66, 56]
    assert walker.torso_idx == 3


def test_task_config_arithmetic(task_cfg):
    assert task_cfg.episode_length == 195                      # train.py:221-225 with the shipped yaml
    assert task_cfg.n_joint_idxs == 33 and task_cfg.n_body_idxs == 18 and task_cfg.n_endeff_idxs == 5
    assert abs(task_cfg.var_coeff - 5e-3) < 1e-9               # yaml overrides the dataclass default (quirk 11)


def test_synthetic_clip_shapes(walker, clips2):
    c = clips2
    assert c.position.shape == (2, 250, 3) and c.quaternion.shape == (2, 250, 4) and c.joints.shape == (2, 250, 67)
    assert c.body_positions.shape == (2, 250, 67, 3) and c.body_quaternions.shape == (2, 250, 67, 4)
    assert c.velocity.shape == (2, 250, 3) and c.angular_velocity.shape == (2, 250, 3) and c.joints_velocity.shape == (2, 250, 67)
    assert np.allclose(np.linalg.norm(c.quaternion, axis=-1), 1, atol=1e-6)
    lo, hi = walker.sections["jnt_range"].reshape(-1, 2)[1:].T
    assert (c.joints >= lo - 1e-6).all() and (c.joints <= hi + 1e-6).all()
    # row 0 is the world body, row 1 the walker root (floor removed): stac-mjx layout
    assert np.allclose(c.body_positions[:, :, 0], 0) and np.allclose(c.body_positions[:, :, 1], c.position, atol=1e-6)
    again = clipmod.make_synthetic_clips(walker.sections, 2)
    assert (again.joints == c.joints).all()


def test_policy_param_layout_matches_c_count():
    """The flattened parameter vector (policy.flatten_params) has exactly the length the C side derives from the descriptor
    (include/tmjx.h, tmjx_policy_param_count), for the reference network (rodent-full-clips.yaml:50-57) and a small one."""
    from track_mjx_b200 import policy as P

    lib = L.load()
    for cfg in (P.IntentionNetworkConfig(), P.IntentionNetworkConfig(obs_size=40, reference_obs_size=24, action_size=3, latent_size=4,
                                                                        encoder_layers=(16, 8), decoder_layers=(8,))):
        p = P.init_params(cfg, seed=0)
        flat = P.flatten_params(cfg, p)
        desc = P.make_desc(cfg)
        assert flat.dtype == np.float32 and flat.size == lib.tmjx_policy_param_count(C.byref(desc))
        # LeCun-uniform bound of the first kernel: sqrt(3 / fan_in)
        k0 = p["encoder/hidden_0/kernel"]
        assert k0.shape == (cfg.reference_obs_size, cfg.encoder_layers[0]) and np.abs(k0).max() <= np.sqrt(3.0 / cfg.reference_obs_size)
    full = P.IntentionNetworkConfig()
    macs = 0
    k = full.reference_obs_size
    for n in full.encoder_layers:
        macs += k * n
        k = n
    macs += k * 2 * full.latent_size
    k = full.latent_size + full.obs_size - full.reference_obs_size
    for n in list(full.decoder_layers) + [2 * full.action_size]:
        macs += k * n
        k = n
    assert abs(2 * macs - 5.48e6) < 0.02e6          # the 5.5 MFLOP / env-step of SURVEY 8f


def test_policy_struct_layout_matches_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "tmjx.h"\nint main(){printf("%zu %zu\\n", sizeof(TmjxPolicyDesc), '
                   '__builtin_offsetof(TmjxPolicyDesc, n_decoder_layers));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(L.PolicyDescC), L.PolicyDescC.n_decoder_layers.offset]


def test_multiclip_loader_round_trip(walker, tmp_path):
    """make_multiclip_data (reference io/load.py:105-137) on stac-mjx-shaped flat arrays reproduces the clip table field by
    field, from a mapping and from an .npz; select_clips / generate_train_test_split partition the clips (load.py:187-278)."""
    clips = clipmod.make_synthetic_clips(walker.sections, 5, clip_length=60)
    flat = clipmod.to_stac_arrays(clips)
    assert flat["qpos"].shape == (300, 74) and flat["qvel"].shape == (300, 73) and flat["xpos"].shape == (300, 67, 3)
    again = clipmod.make_multiclip_data(flat, n_frames_per_clip=60)
    np.savez(tmp_path / "clips.npz", n_frames_per_clip=60, **flat)
    from_file = clipmod.make_multiclip_data(str(tmp_path / "clips.npz"))
    for k in ("position", "quaternion", "joints", "body_positions", "velocity", "angular_velocity", "joints_velocity", "body_quaternions"):
        assert np.array_equal(getattr(again, k), np.asarray(getattr(clips, k), np.float32)), k
        assert np.array_equal(getattr(from_file, k), getattr(again, k)), k
    with pytest.raises(ValueError):
        clipmod.make_multiclip_data(flat, n_frames_per_clip=70)
    train, test = clipmod.generate_train_test_split(again, test_ratio=0.4, rng=np.random.default_rng(0))
    assert train.n_clips == 3 and test.n_clips == 2
    both = np.sort(np.concatenate([train.original_clip_idx[:, 0], test.original_clip_idx[:, 0]]))
    assert np.array_equal(both, np.arange(5))
    assert np.array_equal(test.joints, again.joints[test.original_clip_idx[:, 0]])


def test_bench_reference_arm_prints_one_schema_complete_json_line():
    """`bench.py --impl reference` (the CPU arm: the oracle port on the host cores) prints exactly one JSON line on stdout
    with the keys the measurement contract names."""
    import json
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
