"""Regenerate solorl_b200/assets/{solo8,solo12}.json from the reference URDFs.

Run in the builder container only (needs /root/reference):
    python tools/make_model_assets.py [/root/reference]
The JSON holds numbers only (tree, inertial parameters, joint limits); meshes are
not needed because the foot collision primitive is a sphere (model.py).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from solorl_b200.model import SoloModel, ASSET_DIR  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
for urdf, name in (("solo.urdf", "solo8"), ("solo12.urdf", "solo12")):
    m = SoloModel.from_urdf(os.path.join(ref, "solo_description", "robots", urdf))
    m.name = name
    os.makedirs(ASSET_DIR, exist_ok=True)
    with open(os.path.join(ASSET_DIR, name + ".json"), "w") as f:
        json.dump(m.to_json(), f, indent=1)
    print(name, "links", m.num_links, "nj", m.nj, "mass", m.total_mass,
          "joints_idx", m.joints_idx, "feet_idx", m.feet_idx)
