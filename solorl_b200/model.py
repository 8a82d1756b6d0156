"""Flattened kinematic-tree table for Solo8 / Solo12.

Replaces what the reference obtains from PyBullet at load time
(``p.loadURDF`` + the ``p.getJointInfo`` loop of ``SoloBase.load_parts``,
reference ``solo.py:69-110``): link order = URDF ``<joint>`` order (PyBullet link
index), actuated joints = names without "ANKLE", feet = names with "ANKLE".

The URDF is parsed once on the host into :class:`SoloModel`; the C-ABI receives
it as a ``SoloModelTable`` (``include/solo_b200.h``).  Built-in tables for
``solo.urdf`` / ``solo12.urdf`` ship as JSON under ``assets/`` (numbers extracted
from the reference's ``solo_description/robots/*.urdf`` by
``tools/make_model_assets.py``) so that configs carrying the reference author's
absolute ``model_urdf`` path (``configs/basic.yaml:4``) still resolve.
"""
from __future__ import annotations

import json
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

JOINT_FIXED = 0
JOINT_REVOLUTE = 1

# Foot collision primitive: the foot mesh is a wheel of radius 0.016 m about the
# link y axis centred on the FOOT link origin (measured from
# solo_foot.stl); one sphere per foot stands in for Bullet's convex hull.
DEFAULT_FOOT_RADIUS = 0.016


def _floats(s: Optional[str], n: int, default: float = 0.0) -> List[float]:
    if s is None:
        return [default] * n
    vals = [float(x) for x in s.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} floats, got {s!r}")
    return vals


@dataclass
class SoloModel:
    """Kinematic tree in URDF joint order (see module docstring)."""

    name: str
    link_names: List[str]
    joint_names: List[str]
    parent: List[int]            # -1 = base
    jtype: List[int]
    axis: np.ndarray             # [L,3]
    origin: np.ndarray           # [L,3]
    mass: np.ndarray             # [L]
    com: np.ndarray              # [L,3]
    inertia: np.ndarray          # [L,6] ixx ixy ixz iyy iyz izz
    base_mass: float
    base_com: np.ndarray         # [3]
    base_inertia: np.ndarray     # [6]
    joint_lower: np.ndarray      # [L]
    joint_upper: np.ndarray      # [L]
    foot_radius: float = DEFAULT_FOOT_RADIUS
    foot_center: np.ndarray = field(default_factory=lambda: np.zeros((4, 3)))

    # ---- tables the reference builds in load_parts (solo.py:91-110) ----
    @property
    def num_links(self) -> int:
        return len(self.parent)

    @property
    def joints_idx(self) -> List[int]:
        """PyBullet indices of actuated joints (names without 'ANKLE')."""
        return [i for i, n in enumerate(self.joint_names) if "ANKLE" not in n]

    @property
    def feet_idx(self) -> List[int]:
        return [i for i, n in enumerate(self.joint_names) if "ANKLE" in n]

    @property
    def nj(self) -> int:
        return len(self.joints_idx)

    @property
    def ordered_joint_names(self) -> List[str]:
        return [self.joint_names[i] for i in self.joints_idx]

    @property
    def joint_state_limit(self) -> float:
        """``self.ordered_joints[0].upperLimit`` (solo.py:109)."""
        return float(self.joint_upper[self.joints_idx[0]])

    @property
    def total_mass(self) -> float:
        return float(self.base_mass + self.mass.sum())

    # ---- construction ----
    @staticmethod
    def from_urdf(path: str, foot_radius: float = DEFAULT_FOOT_RADIUS) -> "SoloModel":
        root = ET.parse(path).getroot()
        links = {}
        for ln in root.findall("link"):
            inert = ln.find("inertial")
            if inert is None:
                m, c, I = 0.0, [0.0] * 3, [0.0] * 6
            else:
                o = inert.find("origin")
                if o is not None and any(abs(v) > 0 for v in _floats(o.get("rpy"), 3)):
                    raise ValueError("inertial rpy != 0 is not supported")
                c = _floats(o.get("xyz") if o is not None else None, 3)
                m = float(inert.find("mass").get("value"))
                it = inert.find("inertia")
                I = [float(it.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")]
            links[ln.get("name")] = (m, c, I)

        joints = root.findall("joint")
        children = {j.find("child").get("link") for j in joints}
        bases = [n for n in links if n not in children]
        if len(bases) != 1:
            raise ValueError(f"expected exactly one root link, found {bases}")
        base = bases[0]
        link_index = {base: -1}
        names, jnames, parent, jtype, axis, origin = [], [], [], [], [], []
        mass, com, inertia, lo, hi = [], [], [], [], []
        for j in joints:
            p = j.find("parent").get("link")
            c = j.find("child").get("link")
            if p not in link_index:
                raise ValueError(f"joint {j.get('name')}: parent {p} appears after its child")
            t = j.get("type")
            if t == "revolute" or t == "continuous":
                jt = JOINT_REVOLUTE
            elif t == "fixed":
                jt = JOINT_FIXED
            else:
                raise ValueError(f"joint type {t!r} is not supported")
            o = j.find("origin")
            if o is not None and any(abs(v) > 0 for v in _floats(o.get("rpy"), 3)):
                raise ValueError("joint origin rpy != 0 is not supported")
            ax = j.find("axis")
            lim = j.find("limit")
            link_index[c] = len(names)
            names.append(c)
            jnames.append(j.get("name"))
            parent.append(link_index[p])
            jtype.append(jt)
            axis.append(_floats(ax.get("xyz"), 3) if ax is not None else [1.0, 0.0, 0.0])
            origin.append(_floats(o.get("xyz") if o is not None else None, 3))
            m, cc, I = links[c]
            mass.append(m)
            com.append(cc)
            inertia.append(I)
            lo.append(float(lim.get("lower", 0.0)) if lim is not None else 0.0)
            hi.append(float(lim.get("upper", 0.0)) if lim is not None else 0.0)
        bm, bc, bI = links[base]
        model = SoloModel(
            name=os.path.splitext(os.path.basename(path))[0],
            link_names=names, joint_names=jnames, parent=parent, jtype=jtype,
            axis=np.array(axis, dtype=np.float64), origin=np.array(origin, dtype=np.float64),
            mass=np.array(mass, dtype=np.float64), com=np.array(com, dtype=np.float64),
            inertia=np.array(inertia, dtype=np.float64), base_mass=float(bm),
            base_com=np.array(bc, dtype=np.float64), base_inertia=np.array(bI, dtype=np.float64),
            joint_lower=np.array(lo), joint_upper=np.array(hi), foot_radius=foot_radius)
        model.foot_center = np.zeros((len(model.feet_idx), 3))
        return model

    def to_json(self) -> dict:
        d = {}
        for k, v in self.__dict__.items():
            d[k] = v.tolist() if isinstance(v, np.ndarray) else v
        return d

    @staticmethod
    def from_json(d: dict) -> "SoloModel":
        arr = ("axis", "origin", "mass", "com", "inertia", "base_com", "base_inertia",
               "joint_lower", "joint_upper", "foot_center")
        kw = {k: (np.array(v, dtype=np.float64) if k in arr else v) for k, v in d.items()}
        return SoloModel(**kw)

    @staticmethod
    def builtin(name: str) -> "SoloModel":
        """``'solo8'`` (solo.urdf) or ``'solo12'`` (solo12.urdf)."""
        with open(os.path.join(ASSET_DIR, f"{name}.json")) as f:
            return SoloModel.from_json(json.load(f))

    @staticmethod
    def resolve(model_urdf: str) -> "SoloModel":
        """Resolve the YAML ``model_urdf`` key (baseEnv.py:8): parse the file when it
        exists, otherwise map the basename onto a built-in table."""
        if model_urdf in ("solo8", "solo12"):
            return SoloModel.builtin(model_urdf)
        if os.path.isfile(model_urdf):
            return SoloModel.from_urdf(model_urdf)
        base = os.path.basename(model_urdf)
        if base == "solo.urdf":
            return SoloModel.builtin("solo8")
        if base == "solo12.urdf":
            return SoloModel.builtin("solo12")
        raise FileNotFoundError(f"model_urdf {model_urdf!r} not found and not a built-in name")
