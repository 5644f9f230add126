"""Bird's-eye overview of the map with the car on it — the reference's debugging view (tinycarlo/renderer.py:19-34,53-82,
car.py:171-228), shown by render_mode="human". Not on the step path: a handful of cv2 calls on the host per frame, drawn
from a snapshot of ONE env's state (vector envs pick the env to look at).

The static part (lanelines in their layer colours, the lanepath in grey, optional lanepath node numbers) is drawn once;
render() copies it and adds the chassis outline, the four wheels (front ones turned by the Ackermann angles of the
current steering angle) and the tracked local path."""
import math
from typing import Optional, Sequence, Tuple

import numpy as np

from .maptables import MapTables


def _cv2():
    import cv2  # the reference's own dependency for this view; imported lazily: the step path does not need it
    return cv2


def chassis_points(position, rotation, wheelbase, track_width) -> np.ndarray:
    """corners of the chassis rectangle, rear-left first (car.py:171-181): rear axle centre + R(rot) @ local"""
    c, s = math.cos(rotation), math.sin(rotation)
    local = [(0.0, -track_width / 2), (0.0, track_width / 2), (wheelbase, track_width / 2), (wheelbase, -track_width / 2)]
    return np.array([[position[0] + c * x - s * y, position[1] + s * x + c * y] for x, y in local])


def ackermann_angles(steering_angle_deg: float, wheelbase: float, track_width: float) -> Tuple[float, float]:
    """(front-left, front-right) wheel angles in rad, visual only (car.py:206-221; it works on millimetre-scaled lengths
    and the radius of the last step, which is wheelbase / tan(steering angle), 0 when driving straight, car.py:95-101)"""
    if abs(steering_angle_deg) < 0.0001:
        return 0.0, 0.0
    radius = wheelbase / math.tan(math.radians(steering_angle_deg))
    wb, tw = wheelbase / 1000, track_width / 1000
    inner = -math.atan(wb / (radius - (tw / 2 + 0.000001)))
    outer = -math.atan(wb / (radius + (tw / 2 + 0.000001)))
    return (outer, inner) if radius > 0 else (inner, outer)


def wheel_segments(position, rotation, steering_angle_deg, wheelbase, track_width):
    """four 2-point segments [front-left, front-right, rear-left, rear-right] in world coordinates (car.py:183-204). The
    front wheels turn about their centres; cv2.getRotationMatrix2D(centre, angle_deg, 1) is [[a, b, (1-a)cx - b cy],
    [-b, a, b cx + (1-a) cy]] with a = cos, b = sin."""
    wl = wheelbase / 3
    c, s = math.cos(rotation), math.sin(rotation)

    def world(x, y):
        return [position[0] + c * x - s * y, position[1] + s * x + c * y]

    def turned(pt, centre, ang):
        a, b = math.cos(ang), math.sin(ang)
        x = a * pt[0] + b * pt[1] + (1 - a) * centre[0] - b * centre[1]
        y = -b * pt[0] + a * pt[1] + b * centre[0] + (1 - a) * centre[1]
        return world(x, y)
    fl_a, fr_a = ackermann_angles(steering_angle_deg, wheelbase, track_width)
    hw = track_width / 2
    fl = [turned(p, (wheelbase - wl / 2, -hw), fl_a) for p in ((wheelbase - wl, -hw), (wheelbase, -hw))]
    fr = [turned(p, (wheelbase - wl / 2, hw), fr_a) for p in ((wheelbase - wl, hw), (wheelbase, hw))]
    rl = [world(0.0, -hw), world(wl, -hw)]
    rr = [world(0.0, hw), world(wl, hw)]
    return [np.array(w) for w in (fl, fr, rl, rr)]


class OverviewRenderer:
    def __init__(self, tables: MapTables, overview_pixel_per_meter: int = 266, background_color: Optional[Sequence[int]] = None,
                 line_thickness: int = 1, node_names: bool = False):
        self.tables = tables
        self.ppm = overview_pixel_per_meter
        self.background_color = None if background_color is None else tuple(int(v) for v in background_color)
        self.line_thickness = int(line_thickness)
        self.static = self._static(node_names)

    def _px(self, pts) -> np.ndarray:
        return np.int32(np.array([pts]) * self.ppm)   # metres -> overview pixels, truncated like renderer.py:81-82

    def _static(self, node_names: bool) -> np.ndarray:
        cv2 = _cv2()
        t = self.tables
        h, w = t.dimension
        img = np.zeros((int(h * self.ppm), int(w * self.ppm), 3), np.uint8)
        if self.background_color is not None:
            img[:] = self.background_color
        for c in range(t.n_classes):
            nodes, color = t.class_nodes(c), tuple(int(v) for v in t.colors[c])
            for a, b in t.class_edges(c):
                img = cv2.polylines(img, self._px([nodes[a], nodes[b]]), False, color, self.line_thickness)
        grey = (50, 50, 50)
        if self.background_color is not None and sorted(self.background_color) != [255, 255, 255]:
            grey = (200, 200, 200)
        for a, b in t.lp_edges:
            img = cv2.polylines(img, self._px([t.lp_nodes[a], t.lp_nodes[b]]), False, grey, self.line_thickness)
        if node_names:
            for i, node in enumerate(t.lp_nodes):
                cv2.putText(img, str(i), tuple(int(v) for v in np.int32(np.array(node) * self.ppm)), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (50, 50, 50), 1,
                            cv2.LINE_AA)
        return img

    def render(self, position=None, rotation: float = 0.0, steering_angle_deg: float = 0.0, wheelbase: float = 0.08, track_width: float = 0.03,
               local_path: Sequence[Tuple[int, int]] = ()) -> np.ndarray:
        """position None: the map alone. local_path: lanepath node pairs."""
        cv2 = _cv2()
        img = self.static.copy()
        if position is None:
            return img
        img = cv2.polylines(img, self._px(chassis_points(position, rotation, wheelbase, track_width)), True, (255, 0, 0), self.line_thickness)
        wheel_px = np.int32((wheelbase / 3 / 6) * self.ppm)
        for seg in wheel_segments(position, rotation, steering_angle_deg, wheelbase, track_width):
            img = cv2.polylines(img, self._px(seg), False, (255, 0, 255), int(wheel_px))
        for a, b in local_path:
            img = cv2.polylines(img, self._px([self.tables.lp_nodes[a], self.tables.lp_nodes[b]]), False, (255, 0, 0), self.line_thickness)
        return img
