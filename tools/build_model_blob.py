"""Compile the reference's rodent MJCF into the model-constant table shipped with the package.

    python tools/build_model_blob.py [/root/reference/track_mjx/environment/walker/assets/rodent/rodent.xml]

Runs in the build container (where /root/reference is mounted); the outputs
`track-mjx_b200/assets/rodent_*.tmjx(.json)` are committed so the GPU box needs no reference tree.
The table is derived data (compiled constants as fp32 arrays), not a copy of the XML.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_mjx_b200 import mjcf, model_blob, walker  # noqa: E402

XML = "/root/reference/track_mjx/environment/walker/assets/rodent/rodent.xml"


def main():
    xml = sys.argv[1] if len(sys.argv) > 1 else XML
    out_dir = os.path.join(os.path.dirname(os.path.abspath(walker.__file__)), "assets")
    os.makedirs(out_dir, exist_ok=True)
    for torque, scale in ((True, 0.9),):
        model = mjcf.compile_mjcf(xml, torque_actuators=torque, rescale_factor=scale)
        path = os.path.join(out_dir, walker.blob_name(torque, scale))
        with open(path, "wb") as f:
            f.write(model_blob.pack(model))
        with open(path + ".json", "w") as f:
            json.dump(dict(body=model["body_names"], joint=model["jnt_names"], actuator=model["actuator_names"],
                           total_mass=float(model["body_mass"].sum()), source=os.path.basename(xml),
                           torque_actuators=torque, rescale_factor=scale), f, indent=1)
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
