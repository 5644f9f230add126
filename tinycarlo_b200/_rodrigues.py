"""Axis-angle -> rotation matrix. Uses cv2.Rodrigues (the reference's own call, camera.py:152-153) when OpenCV is
importable — it is a hard dependency of the reference (setup.py:19) — and a textbook formula otherwise."""
import numpy as np

try:
    import cv2 as _cv2
except Exception:  # pragma: no cover - cv2 is present wherever the reference runs
    _cv2 = None


def rodrigues(rvec) -> np.ndarray:
    rvec = np.asarray(rvec, np.float64).reshape(3)
    if _cv2 is not None:
        return _cv2.Rodrigues(rvec)[0]
    theta = float(np.linalg.norm(rvec))
    if theta < 1e-300:
        return np.eye(3)
    k = rvec / theta
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(theta) * np.eye(3) + (1 - np.cos(theta)) * np.outer(k, k) + np.sin(theta) * Kx
