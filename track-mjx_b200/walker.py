"""Rodent walker: compiled model constants + name -> id tables.

Mirror of reference `track_mjx/environment/walker/rodent.py:16-114` (`Rodent`) and of the index
properties of `walker/base.py:69-135` (`BaseWalker`).  The reference compiles
`assets/rodent/rodent.xml` with the MuJoCo C library at construction; this build ships the model
as a pre-compiled constant table (`assets/rodent_torque_s0.9.tmjx` + `.json` name tables) produced by
`tools/build_model_blob.py` with the MJCF-subset compiler in `mjcf.py`, so that nothing from the
reference tree is needed at run time.  Passing `xml_path=` recompiles from an MJCF file instead.
"""

from __future__ import annotations

import json
import os
from typing import Sequence

import numpy as np

from . import config as _config
from . import model_blob

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def blob_name(torque_actuators: bool, rescale_factor: float) -> str:
    return f"rodent_{'torque' if torque_actuators else 'position'}_s{rescale_factor:g}.tmjx"


class Rodent:
    """Rodent walker (reference walker/rodent.py:16)."""

    def __init__(
        self,
        joint_names: Sequence[str] = _config.RODENT_JOINT_NAMES,
        body_names: Sequence[str] = _config.RODENT_BODY_NAMES,
        end_eff_names: Sequence[str] = _config.RODENT_END_EFF_NAMES,
        *,
        torque_actuators: bool = False,
        rescale_factor: float = 0.9,
        xml_path: str | None = None,
    ):
        self._torso_name = "torso"
        self._joint_names = list(joint_names)
        self._body_names = list(body_names)
        self._end_eff_names = list(end_eff_names)
        if xml_path is not None:
            from . import mjcf

            model = mjcf.compile_mjcf(xml_path, torque_actuators=torque_actuators, rescale_factor=rescale_factor)
            self.blob = model_blob.pack(model)
            self._names = dict(body=model["body_names"], joint=model["jnt_names"], actuator=model["actuator_names"])
        else:
            path = os.path.join(_ASSETS, blob_name(torque_actuators, rescale_factor))
            if not os.path.exists(path):
                raise FileNotFoundError(
                    f"no pre-compiled model {path}; run tools/build_model_blob.py or pass xml_path=")
            with open(path, "rb") as f:
                self.blob = f.read()
            with open(path + ".json") as f:
                self._names = json.load(f)
        self.sections = model_blob.unpack(self.blob)
        d = self.sections["dims"]
        self.nq, self.nv, self.nu, self.na, self.nbody, self.njnt = (int(x) for x in d[:6])
        self.ncon, self.nefc = int(d[7]), int(d[8])
        self._initialize_indices()

    # mj_name2id equivalents
    def body_id(self, name: str) -> int:
        return self._names["body"].index(name)

    def joint_id(self, name: str) -> int:
        return self._names["joint"].index(name)

    def _initialize_indices(self) -> None:
        """reference walker/rodent.py:89-114."""
        self._joint_idxs = np.array([self.joint_id(j) for j in self._joint_names], np.int32)
        self._body_idxs = np.array([self.body_id(b) for b in self._body_names], np.int32)
        self._endeff_idxs = np.array([self.body_id(e) for e in self._end_eff_names], np.int32)
        self._torso_idx = self.body_id(self._torso_name)

    joint_idxs = property(lambda self: self._joint_idxs)
    body_idxs = property(lambda self: self._body_idxs)
    endeff_idxs = property(lambda self: self._endeff_idxs)
    torso_idx = property(lambda self: self._torso_idx)
