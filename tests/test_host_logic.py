"""CPU tests of the host-side logic: config defaults, map tables, camera matrices, spawn draws, C-ABI symbols."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, Golden
from tinycarlo_b200 import _lib
from tinycarlo_b200.camera_params import camera_row, extrinsic_matrix, intrinsic_matrix
from tinycarlo_b200.config import camera_params, car_param_row, load_config, resolve_map_path, sim_params
from tinycarlo_b200.maptables import MapTables
from tinycarlo_b200.spawn import SpawnSampler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPAWN_KNUFF = [156, 18, 217, 214, 325, 354, 176, 402, 339, 376, 385, 419, 396, 37, 149, 62, 240, 113, 98, 299, 2]
SPAWN_SIMPLE = [57, 143, 112, 121, 138, 157, 67, 46, 165, 124, 79, 33, 84, 21, 178, 7]


def test_library_exports_every_declared_symbol():
    """include/tinycarlo_b200.h <-> libtinycarlo_b200.so (no compute calls: there is no GPU here)."""
    hdr = open(os.path.join(ROOT, "include", "tinycarlo_b200.h")).read()
    declared = set(re.findall(r"TC_API\s+[\w\s\*]+?\b(tc_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert L.tc_abi_version() == 1


def test_product_has_no_cpu_path():
    """Without a CUDA device the env refuses to construct (no fallback); the product never imports oracle/."""
    import torch
    if not torch.cuda.is_available():
        from tinycarlo_b200 import TinyCarloVecEnv, TinyCarloError
        with pytest.raises(TinyCarloError):
            TinyCarloVecEnv({"sim": {}, "car": {}, "camera": {"max_range": 0.5}, "map": {"map_name": "simple_layout", "pixel_per_meter": 450}}, 2)
    pkg = os.path.join(ROOT, "tinycarlo_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):   # the Python product never imports the oracle nor loads the host test build
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "tc_oracle" not in src and "hosttest" not in src, fn


def test_config_defaults_and_errors(tmp_path):
    cfg = {"sim": {}, "car": {}, "camera": {"max_range": 0.5}, "map": {"json_path": "m.json", "pixel_per_meter": 100}}
    c, path = load_config(cfg)
    assert path is None and sim_params(c) == {"fps": 30, "T": 1 / 30, "observation_space_format": "rgb", "render_realtime": False,
                                                "overview_pixel_per_meter": 150, "render_node_names": False}
    row = car_param_row(c["car"], 1 / 30)
    assert row[:4] == [0.08, 0.03, 1.0, 35.0] and all(math.isnan(v) for v in row[4:7]) and row[7] == 1 / 30
    cam = camera_params(c["camera"])
    assert cam["resolution"] == [128, 160] and cam["fov"] == 90 and cam["line_thickness"] == 1 and cam["position"] == [0, 0, 0]
    with pytest.raises(KeyError):
        load_config({"sim": {}, "car": {}, "camera": {}})           # missing section -> KeyError like the reference
    with pytest.raises(ValueError):
        camera_params({})                                            # max_range None: the reference crashes on the first frame
    with pytest.raises(TypeError):
        car_param_row({"max_acceleration": 0.1}, 1 / 30)            # car.py:82 `None * dt`
    # yaml path and directory forms; json_path relative to the yaml (map.py:15-16)
    (tmp_path / "config.yaml").write_text("sim: {fps: 20}\ncar: {}\ncamera: {max_range: 1.0}\nmap: {json_path: maps/x.json, pixel_per_meter: 10}\n")
    for arg in (str(tmp_path / "config.yaml"), str(tmp_path)):
        c, path = load_config(arg)
        assert c["sim"]["fps"] == 20 and path == str(tmp_path / "config.yaml")
        assert resolve_map_path(c["map"], path) == str(tmp_path / "maps/x.json")
    assert resolve_map_path({"json_path": "a/b.json"}, None) == "./a/b.json"


def test_map_tables_follow_the_reference_loader():
    t = MapTables(resolve_map_path({"map_name": "knuffingen"}, None), 222, SPAWN_KNUFF)
    assert t.class_names == ["outer", "dashed", "solid", "hold", "area"]
    assert list(np.diff(t.ll_node_off)) == [517, 111, 149, 18, 32] and list(np.diff(t.ll_edge_off)) == [499, 81, 116, 11, 32]
    assert t.lp_nodes.shape == (429, 2) and t.lp_edges.shape == (443, 2)
    import json
    raw = json.load(open(resolve_map_path({"map_name": "knuffingen"}, None)))
    n0 = raw["lanepath"]["nodes"][113]
    assert t.lp_nodes[113][0] == n0[0] / 222 and t.lp_nodes[113][1] == n0[1] / 222   # SURVEY Appendix B spawn node
    assert list(t.lp_nodes[113]) == [4.202702702702703, 0.8153153153153153]
    e = np.nonzero(t.lp_edges[:, 0] == 113)[0][0]
    assert t.lp_orient[e] == -0.3268266529299409
    assert np.allclose(np.abs(np.abs(t.lp_orient - t.lp_orient_rev) - math.pi), 0, atol=1e-12)
    assert int((~t.has_successor).sum()) == 6


def test_camera_matrices_match_reference_values():
    """E and K of the shipped Knuffingen camera at 480x640, recorded from the live reference."""
    E = extrinsic_matrix([0.0, -0.005, 0.04], [22, 0, 0])
    want = [[6.123233995736766e-17, -1, 0, -0.005], [0.37460659341591196, 2.2938038278314527e-17, 0.9271838545667874, -0.0370873541826715],
            [-0.9271838545667874, -5.677363698581607e-17, 0.37460659341591196, -0.01498426373663648]]
    np.testing.assert_allclose(E, np.array(want), rtol=1e-15, atol=1e-17)
    assert np.array_equal(E, Golden("knuff_480_stanley")["E"][0])   # bit for bit what the reference computed
    K = intrinsic_matrix(80, [480, 640])
    assert K[0, 0] == 381.3611496301472 and K[1, 1] == 286.02086222261045 and K[0, 2] == 320 and K[1, 2] == 240
    for name in ("knuff_camrand",):
        g = Golden(name)
        muts = {int(k): v for k, v in g.meta["cam_mutations"].items()}
        cc = dict(g.cfg["camera"])
        for f in range(g.F):
            t = int(g["ev_step"][f])
            if g["ev_kind"][f] == 1 and t in muts:
                cc.update(muts[t])
            row = camera_row(cc["position"], cc["orientation"], cc["fov"], cc["resolution"], cc["max_range"])
            if g["ev_kind"][f] == 1:
                assert np.array_equal(row[:12].reshape(3, 4), g["E"][f]), f
                assert row[12] == g["K"][f][0][0] and row[13] == g["K"][f][1][1]


def test_spawn_sampler_matches_reference_draws():
    d = np.load(os.path.join(GOLDEN_DIR, "spawn_draws.npz"))
    for key, (mname, ppm, sp) in {"knuffingen_default": ("knuffingen", 222, SPAWN_KNUFF), "knuffingen_none": ("knuffingen", 222, None),
                                  "simple_layout_default": ("simple_layout", 450, SPAWN_SIMPLE), "simple_layout_none": ("simple_layout", 450, None)}.items():
        t = MapTables(resolve_map_path({"map_name": mname}, None), ppm, sp)
        want = d[key]                      # [64 seeds, 12 consecutive resets]
        # a vector env seeded with s gives env i the reference's stream for seed s + i
        s = SpawnSampler(t, 16, table_len=5, env_index_offset=3)
        tab = s.seed(10)
        assert np.array_equal(tab, want[13:29, :5]), key
        # consuming entries and refilling continues each env's stream
        consumed = np.array([i % 4 for i in range(16)])
        tab = s.advance(consumed).copy()
        for i in range(16):
            assert np.array_equal(tab[i], want[13 + i, consumed[i]:consumed[i] + 5]), (key, i)


def test_vectorised_pcg64_equals_numpy_generator():
    """tinycarlo_b200/pcg64.py against numpy itself: SeedSequence pool, raw PCG64 outputs, buffered 32-bit halves and
    Lemire-bounded integers (incl. ranges with ~50 % rejection), for small and multi-word seeds, with masks."""
    from tinycarlo_b200.pcg64 import VecPCG64, seed_sequence_state
    seeds = np.array(list(range(0, 200)) + [2**32 - 1, 2**32, 2**32 + 5, 2**40 + 123, 2**63 + 7, 2**64 - 1, 123456789012], dtype=np.uint64)
    ss = seed_sequence_state(seeds)
    for i, s in enumerate(seeds):
        assert np.array_equal(ss[i], np.random.SeedSequence(int(s)).generate_state(4, np.uint64)), int(s)
    v = VecPCG64(seeds)
    bg = [np.random.PCG64(np.random.SeedSequence(int(s))) for s in seeds]
    everyone = np.ones(len(seeds), bool)
    for _ in range(4):
        assert np.array_equal(v.next_uint64(everyone), np.array([b.random_raw() for b in bg], dtype=np.uint64))
    v = VecPCG64(seeds)
    gens = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(s)))) for s in seeds]
    for hi in [21, 16, 429, 428, 2**31 + 1, 3, 2**32, 1, 7, 100000, 21]:
        assert np.array_equal(v.bounded(hi), np.array([int(g.integers(0, hi)) for g in gens])), hi
    lst = list(range(100, 121))
    mask = np.arange(len(seeds)) % 3 == 0
    got = np.array(lst)[v.bounded(len(lst), mask)]
    assert np.array_equal(got, np.array([int(g.choice(lst)) for g, m in zip(gens, mask) if m]))
    assert np.array_equal(v.bounded(428), np.array([int(g.integers(0, 428, size=1, dtype=int)[0]) for g in gens]))
    # checkpoint round trip
    st = v.state_dict()
    a = v.bounded(1000)
    w = VecPCG64()
    w.load_state_dict(st)
    assert np.array_equal(w.bounded(1000), a)


def test_headline_kernel_keeps_its_register_budget():
    """the per-class render kernel (u8) sits at 64 registers without spills; local-memory traffic behind
    the observation stores costs it the HBM write roofline. Checked on the built library (no GPU needed)."""
    import re
    import shutil
    import subprocess
    from tinycarlo_b200 import _lib
    if shutil.which("cuobjdump") is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    m = re.search(r"Function _Z24tc_render_classes_kernelILi256ELi0EEv12TcRenderArgs:\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m, "headline kernel not found in the library"
    assert int(m.group(1)) <= 64 and int(m.group(2)) == 0, m.group(0)
    m = re.search(r"Function _Z20tc_render_env_kernelILi256ELi0EEv15TcRenderEnvArgs:\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m and int(m.group(1)) <= 64, "block-per-env kernel missing or over its register budget"


@pytest.mark.parametrize("config", [3, 1, 4, 5])
def test_bench_reference_arm_prints_the_contract_line(config):
    """bench.py --impl reference (the CPU arm: the unmodified reference from baseline/_ref, one process per host core, plus
    the oracle port as a second figure; the port alone when baseline/_ref is absent) prints ONE JSON line with the keys the
    driver reads, and the same `config` object as this repo's arm would."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", str(config), "--steps", "2", "--warmup", "1",
                        "--cpu-envs", "24", "--cpu-seconds", "1"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 2 and "workload" in d["config"] and d["config"]["baseline_config"] == config
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    have_ref = os.path.isdir(os.path.join(root, "baseline", "_ref", "tinycarlo"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    if have_ref:
        assert d["cpu_baseline"]["single_process"]["value"] > 0 and d["cpu_baseline_port"]["kind"] == "port" and d["cpu_baseline_port"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    if config == 3:
        assert d["metric"] == "env-steps/sec (480x640 class obs, Knuffingen)"     # BASELINE.json's metric, unchanged
        sys.path.insert(0, os.path.join(root, "baseline"))
        sys.path.insert(0, root)
        import bench
        import workloads as WL

        class A:
            config, envs_per_gpu = 3, 0
        assert d["config"] == bench.config_dict(WL.WORKLOADS[3], A, 1)      # what the CUDA arm prints: the driver's same_config check


def test_shard_groups_partition_every_group():
    from tinycarlo_b200.distributed import shard_groups
    sizes = [10922, 10922, 10924]
    for world in (1, 2, 4, 8, 3):
        seen = [np.zeros(n, int) for n in sizes]
        base = np.cumsum([0] + sizes)
        for r in range(world):
            local, offs = shard_groups(sizes, r, world)
            for g, (n, o) in enumerate(zip(local, offs)):
                seen[g][o - base[g]:o - base[g] + n] += 1
        assert all((s == 1).all() for s in seen), world


def test_gymnasium_registration_with_a_gymnasium_on_the_path():
    """import tinycarlo_b200 registers tinycarlo-v2 when a gymnasium is importable (tinycarlo/__init__.py:3); checked in a fresh
    interpreter with the stand-in of tests/golden/gym_stub (gymnasium itself is not in the image). No GPU needed."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import gymnasium as gym, tinycarlo_b200\n"
            "from tinycarlo_b200 import gym_compat\n"
            "assert gym_compat.HAVE_GYMNASIUM and gym_compat.Env is gym.Env\n"
            "assert gym.envs.registration.registry['tinycarlo-v2'] == 'tinycarlo_b200.env:TinyCarloEnv'\nprint('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(root, "tests", "golden", "gym_stub"), root]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_bench_product_arm_is_free_of_test_infrastructure():
    """the product arm of bench.py must not build its workload from tests/ or load the oracle. bench.py imports neither at
    module level, never puts tests/ on the path, and only its CPU arm (port_arm) names the oracle package."""
    import ast
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "bench.py")).read()
    assert "pair_util" not in src and '"tests"' not in src
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name != "port_arm":
            for sub in ast.walk(node):
                if isinstance(sub, (ast.Import, ast.ImportFrom)):
                    names = [a.name for a in sub.names] + [getattr(sub, "module", "") or ""]
                    assert not any(n.split(".")[0] == "oracle" for n in names), (node.name, names)
    out = subprocess.run([sys.executable, "-c", "import sys, bench; print(sorted(m for m in sys.modules if m.split('.')[0] in ('oracle', 'pair_util', 'golden_util')))"],
                         cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "[]", (out.stdout, out.stderr[-500:])
    for f in os.listdir(os.path.join(root, "tools")):     # the tools parse
        if f.endswith(".py"):
            ast.parse(open(os.path.join(root, "tools", f)).read())
