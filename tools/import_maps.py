"""Re-serialise the reference's public map datasets into tinycarlo_b200/maps/ (compact JSON).

The maps (node/edge graphs in pixels) are INPUT DATA of the benchmark configs named in
BASELINE.json (simple_layout, Knuffingen), not source code; the schema is the mapbuilder's
(reference mapbuilder/mapbuilder.py:93-98): {width,height,lanelines:{name:{layer_color,nodes,edges}},lanepath:{...}}.
Key order inside "lanelines" is the class order and is preserved.
Run once in the build container:  python tools/import_maps.py
"""
import json
import os
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples/maps"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tinycarlo_b200", "maps")

for name in ("simple_layout", "knuffingen", "formula_student_track", "formula_student_skidpad"):
    with open(os.path.join(SRC, name + ".json")) as f:
        data = json.load(f)
    with open(os.path.join(DST, name + ".json"), "w") as f:
        json.dump(data, f, separators=(",", ":"))
    print(name, os.path.getsize(os.path.join(DST, name + ".json")), "bytes")
