"""Helpers to read the golden traces in tests/golden/ (recorded from the unmodified reference)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SCENARIOS = ["simple_rgb_smoke", "knuff_480_stanley", "knuff_mixed", "simple_84_random", "simple_480_rgb_random",
             "knuff_thick1", "knuff_thick3", "knuff_thick6", "knuff_camrand", "fs_track_free", "fs_skidpad_rgb",
             "knuff_wrapped_cte", "knuff_wrapped_lane"]


class Golden:
    def __init__(self, name):
        self.name = name
        with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
            self.d = {k: z[k] for k in z.files}   # decompress once (NpzFile re-reads the zip member on every access)
        self.meta = json.loads(str(self.d["meta"]))
        self.cfg = self.meta["config"]
        self.F = len(self.d["ev_kind"])
        self.fmt = self.cfg["sim"]["observation_space_format"]
        self.H, self.W = self.cfg["camera"]["resolution"]
        self.class_names = self.meta["class_names"]
        self.C = len(self.class_names)
        self.wrapped = len(self.meta["wrappers"]) > 0

    def __getitem__(self, k):
        return self.d[k]

    def max_range_per_frame(self):
        """max_range in force when frame f was rendered (camera mutations are applied before step t)."""
        mr = float(self.cfg["camera"]["max_range"])
        muts = {int(k): v for k, v in self.meta.get("cam_mutations", {}).items()}
        out = np.zeros(self.F)
        kind, step = self.d["ev_kind"], self.d["ev_step"]
        cur = mr
        applied = set()
        for f in range(self.F):
            t = int(step[f])
            if kind[f] == 1 and t in muts and t not in applied:
                cur = float(muts[t].get("max_range", cur))
                applied.add(t)
            out[f] = cur
        return out

    def segments(self, f):
        """per-class list of int32 [k,4] arrays for frame f"""
        off = self.d["seg_off"][f]
        return [self.d["seg_i32"][off[c]:off[c + 1]] for c in range(self.C)], \
               [self.d["seg_f64"][off[c]:off[c + 1]] for c in range(self.C)]

    def classes_frame(self, f):
        bits = np.unpackbits(self.d["cls_bits"][f])[: self.C * self.H * self.W]
        return (bits.reshape(self.C, self.H, self.W) * 255).astype(np.uint8)
