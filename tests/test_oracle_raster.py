"""Fuzzes the oracle's restatement of cv2.polylines (oracle/tc_oracle.c) against the
installed OpenCV (opencv-python-headless 4.13.0). cv2 is third-party arithmetic of the reference path
(renderer.py:43,50); it exists in this image on the CPU box and on the GPU box, so this runs everywhere. CPU-only."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import oracle as orc  # noqa: E402

INT_MIN = -2**31


def check(H, W, p0, p1, t, nch=1, color=255):
    shape = (H, W) if nch == 1 else (H, W, nch)
    a = np.zeros(shape, np.uint8)
    b = np.zeros(shape, np.uint8)
    cv2.polylines(a, np.int32([[p0, p1]]), False, color, t)
    orc.polyline(b, p0, p1, color if nch > 1 else [color], t)
    return np.array_equal(a, b)


def test_cv2_version_is_the_pinned_one():
    assert cv2.__version__.startswith("4.13."), "rasteriser parity is pinned to OpenCV 4.13.0 (SURVEY H1)"


@pytest.mark.parametrize("H,W,spread,n", [(24, 32, 12, 6000), (48, 64, 300, 3000), (84, 84, 40, 2000), (480, 640, 200, 150)])
def test_near_frame(H, W, spread, n):
    rng = np.random.default_rng(H * 1000 + W)
    bad = []
    for _ in range(n):
        t = int(rng.integers(1, 9))
        p0 = (int(rng.integers(-spread, W + spread)), int(rng.integers(-spread, H + spread)))
        p1 = (int(rng.integers(-spread, W + spread)), int(rng.integers(-spread, H + spread)))
        if not check(H, W, p0, p1, t):
            bad.append((p0, p1, t))
    assert not bad, bad[:5]


@pytest.mark.parametrize("mag", [10**4, 10**6, 10**9, 2**31 - 1])
def test_one_endpoint_far_away(mag):
    """near-plane fix-ups send endpoints to |coord| ~ 1e9: clipLine in double on int64."""
    rng = np.random.default_rng(mag % 9973)
    H, W = 48, 64
    bad = []
    for _ in range(2500):
        t = int(rng.integers(1, 7))
        p0 = (int(rng.integers(-5, W + 5)), int(rng.integers(-5, H + 5)))
        p1 = (int(rng.integers(-mag, mag + 1)), int(rng.integers(-mag, mag + 1)))
        if rng.random() < 0.5:
            p0, p1 = p1, p0
        if not check(H, W, p0, p1, t):
            bad.append((p0, p1, t))
    assert not bad, bad[:5]


def test_int_min_endpoints():
    """np.int32 of NaN/inf/overflow is INT_MIN; OpenCV draws those and so must we."""
    H, W = 48, 64
    for t in (1, 2, 3, 6):
        for p0, p1 in [((10, 10), (INT_MIN, INT_MIN)), ((INT_MIN, 5), (20, 20)), ((INT_MIN, INT_MIN), (INT_MIN, INT_MIN)),
                       ((30, INT_MIN), (30, 40)), ((INT_MIN, 20), (2**31 - 1, 20)), ((5, 5), (5, 5))]:
            assert check(H, W, p0, p1, t), (p0, p1, t)


def test_degenerate_and_shapes():
    H, W = 24, 32
    for t in range(1, 9):
        assert check(H, W, (10, 10), (10, 10), t)       # zero length: caps only
        assert check(H, W, (3, 7), (10, 7), t)          # horizontal
        assert check(H, W, (7, 3), (7, 15), t)          # vertical
        assert check(H, W, (0, 0), (W - 1, H - 1), t)   # corner to corner
        assert check(H, W, (-3, -3), (W + 3, H + 3), t)
    img = np.zeros((H, W), np.uint8)
    orc.polyline(img, (10, 10), (10, 10), [255], 2)
    assert int((img > 0).sum()) == 5                    # plus sign
    img[:] = 0
    orc.polyline(img, (10, 10), (10, 10), [255], 3)
    assert int((img > 0).sum()) == 13                   # radius-2 disc


def test_rgb_painters_order():
    """renderer.py:41-43: later layers overwrite earlier ones, 3 channels."""
    rng = np.random.default_rng(7)
    H, W = 40, 56
    for _ in range(200):
        a = np.zeros((H, W, 3), np.uint8)
        b = np.zeros((H, W, 3), np.uint8)
        for _layer in range(4):
            col = [int(c) for c in rng.integers(0, 256, 3)]
            t = int(rng.integers(1, 5))
            for _seg in range(3):
                p0 = (int(rng.integers(-20, W + 20)), int(rng.integers(-20, H + 20)))
                p1 = (int(rng.integers(-20, W + 20)), int(rng.integers(-20, H + 20)))
                cv2.polylines(a, np.int32([[p0, p1]]), False, col, t)
                orc.polyline(b, p0, p1, col, t)
        assert np.array_equal(a, b)
