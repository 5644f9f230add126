#!/usr/bin/env python
"""Installs the UNMODIFIED reference into baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).

    python baseline/install_ref.py            # needs /root/reference (the build container); no-op message elsewhere

What lands in baseline/_ref/:
  tinycarlo/ + tinycarlo-2.0.0.dist-info/   `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of
                                            /root/reference>` (from a copy under /tmp: the checkout is read-only and setuptools writes
                                            build/ and *.egg-info into the source tree; --no-deps because gymnasium is not in the image)
  examples/maps/*.json, examples/*.yaml     the map and config files the reference's examples load (data, copied as they are)
  gymnasium/                                the ~100-line stand-in of tests/golden/gym_stub (gymnasium is absent from the image and
                                            there is no network): register / make / Env seeding as gymnasium >= 0.26 / Wrapper / spaces
Nothing of this is product code and nothing in tinycarlo_b200/ imports it: bench.py --impl reference (baseline/ref_arm.py) times it."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("TINYCARLO_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def install(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(REF, "tinycarlo")):
        if verbose:
            print(f"install_ref: {REF} not present, keeping baseline/_ref as it is ({'present' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(os.path.join(DST, "tinycarlo"))
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "models", "docs", "*.npy", "data"))
        cmd = [sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DST, src]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    os.makedirs(os.path.join(DST, "examples", "maps"))
    for f in os.listdir(os.path.join(REF, "examples", "maps")):
        if f.endswith(".json"):
            shutil.copy(os.path.join(REF, "examples", "maps", f), os.path.join(DST, "examples", "maps", f))
    for f in ("config_knuffingen.yaml", "config_simple_layout.yaml"):
        shutil.copy(os.path.join(REF, "examples", f), os.path.join(DST, "examples", f))
    shutil.copytree(os.path.join(ROOT, "tests", "golden", "gym_stub", "gymnasium"), os.path.join(DST, "gymnasium"),
                    ignore=shutil.ignore_patterns("__pycache__"))
    # the installed package must be byte-identical to the checkout
    for dirpath, _, files in os.walk(os.path.join(REF, "tinycarlo")):
        for f in files:
            if not f.endswith(".py"):
                continue
            a = os.path.join(dirpath, f)
            b = os.path.join(DST, os.path.relpath(a, REF))
            if os.path.exists(b):   # setup.py only packages tinycarlo and tinycarlo.wrapper
                assert open(a, "rb").read() == open(b, "rb").read(), f"{b} differs from the reference"
    if verbose:
        print("install_ref: baseline/_ref ready")
    return True


if __name__ == "__main__":
    install()
