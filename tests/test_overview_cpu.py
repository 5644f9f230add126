"""The bird's-eye overview (tinycarlo_b200/overview.py) against frames recorded from the reference's Renderer.render_overview()
(tests/golden/gen_overview.py), and the map-file validation of MapTables."""
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR
from tinycarlo_b200.config import resolve_map_path
from tinycarlo_b200.maptables import MapTables
from tinycarlo_b200.overview import OverviewRenderer

cv2 = pytest.importorskip("cv2")


def test_overview_matches_reference_frames():
    with np.load(os.path.join(GOLDEN_DIR, "overview.npz")) as z:
        d = {k: z[k] for k in z.files}
    meta = json.loads(str(d["meta"]))
    for k, m in enumerate(meta):
        cfg = m["config"]
        t = MapTables(resolve_map_path(cfg["map"], None), cfg["map"]["pixel_per_meter"], cfg["map"].get("spawn_points"))
        r = OverviewRenderer(t, m["overview_pixel_per_meter"], m["background_color"], m["line_thickness"], m["node_names"])
        assert np.array_equal(r.render(None), d[f"static_{k}"]), ("static", k)
        shape = tuple(d[f"shape_{k}"])
        want_bits = np.unpackbits(d[f"frames_{k}"])[:int(np.prod(shape))].reshape(shape).astype(bool)
        for i, st in enumerate(d[f"states_{k}"]):
            lp = [tuple(int(v) for v in e) for e in d[f"lp_{k}_{i}"] if e[0] >= 0]
            img = r.render([st[0], st[1]], st[2], st[3], cfg["car"]["wheelbase"], cfg["car"]["track_width"], lp)
            assert img.shape == shape[1:]
            assert np.array_equal(img > 0, want_bits[i]), (k, i)
            assert int(img.astype(np.int64).sum()) == int(d[f"sums_{k}"][i]), (k, i)


def _good():
    return {"height": 100, "width": 200, "lanelines": {"solid": {"layer_color": [255, 255, 255], "nodes": [[0, 0], [10, 0], [20, 5]], "edges": [[0, 1], [1, 2]]}},
            "lanepath": {"layer_color": [0, 0, 0], "nodes": [[0, 3], [10, 3], [20, 8]], "edges": [[0, 1], [1, 2]]}}


def test_map_validation_accepts_a_good_map_and_names_what_is_wrong():
    t = MapTables(_good(), 100, [0, 1])
    assert t.n_classes == 1 and len(t.lp_edges) == 2
    cases = []
    m = _good(); del m["lanepath"]; cases.append((m, None, "lanepath"))
    m = _good(); m["lanelines"] = {}; cases.append((m, None, "laneline"))
    m = _good(); m["lanelines"]["solid"]["edges"].append([1, 7]); cases.append((m, None, "edge"))
    m = _good(); m["lanepath"]["edges"] = [[0, 3]]; cases.append((m, None, "edge"))
    m = _good(); m["lanelines"]["solid"]["nodes"][1] = [float("nan"), 0]; cases.append((m, None, "finite"))
    m = _good(); m["lanelines"]["solid"]["layer_color"] = [255, 255]; cases.append((m, None, "layer_color"))
    m = _good(); m["lanepath"]["edges"] = []; cases.append((m, None, "lanepath"))
    cases.append((_good(), [5], "spawn"))          # not a lanepath node
    cases.append((_good(), [2], "spawn"))          # node without successor: the reference would redraw forever
    for data, spawn, word in cases:
        with pytest.raises(ValueError) as e:
            MapTables(data, 100, spawn)
        assert word in str(e.value), (word, str(e.value))
    with pytest.raises(ValueError):
        MapTables(_good(), 0)


def test_shipped_maps_validate():
    for name, ppm in (("knuffingen", 222), ("simple_layout", 450), ("formula_student_track", 300), ("formula_student_skidpad", 200)):
        t = MapTables(resolve_map_path({"map_name": name}, None), ppm)
        assert t.n_classes >= 1 and t.has_successor.any()
