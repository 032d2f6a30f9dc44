"""Host-side data formats on either side of the SRUKF path.

* the odometry text file the reference reads one line per frame (CSLAM::getOneMomentData, SLAM.cpp:462-496),
* the odometry -> control conversion at the head of predictMotion (SLAM.cpp:1444-1454),
* the RobotPath.txt trajectory record (CSLAM::recordRobotInformation, SLAM.cpp:3512-3562).

Pure host code: it feeds `Ut` to CSLAMBatch.predictMotion and stores what comes back; nothing here computes on the
filter state.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import IO, Iterable

import numpy as np

# "%d : %*lf %lf %lf %lf" (SLAM.cpp:475): image id, a skipped number, then x, y, theta.
_NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|nan)"
_LINE = re.compile(rf"\s*([-+]?\d+)\s*:\s*{_NUM}\s+({_NUM})\s+({_NUM})\s+({_NUM})", re.IGNORECASE)


def parse_odometry_line(line: str):
    """One line of the odometry file -> (id, x, y, theta), or None when sscanf would not fill all four fields."""
    m = _LINE.match(line)
    if m is None:
        return None
    return int(m.group(1)), float(m.group(2)), float(m.group(3)), float(m.group(4))


@dataclass
class OdometryTrack:
    """m_odoXY / m_odoTheta of the reference: odometry re-based onto the filter's initial position."""
    init_pos: tuple[float, float] = (0.0, 0.0)      # m_initPos = (m_X_k[0], m_X_k[1]) at counter 0 (SLAM.cpp:485-486)
    ids: list[int] = field(default_factory=list)
    xy: list[tuple[float, float]] = field(default_factory=list)
    theta: list[float] = field(default_factory=list)
    _init_odo: tuple[float, float] | None = None

    def push(self, id_: int, x: float, y: float, theta: float):
        """SLAM.cpp:477-495: the first sample defines the offset, later ones are init_pos + (odo - init_odo)."""
        if self._init_odo is None:
            self._init_odo = (x, y)
            self.xy.append((self.init_pos[0], self.init_pos[1]))
        else:
            self.xy.append((self.init_pos[0] + (x - self._init_odo[0]), self.init_pos[1] + (y - self._init_odo[1])))
        self.ids.append(id_)
        self.theta.append(theta)

    def __len__(self):
        return len(self.ids)

    def control(self, counter: int) -> np.ndarray:
        """Ut = (rot1, trans, rot2) between samples counter-1 and counter (SLAM.cpp:1444-1454)."""
        if not 1 <= counter < len(self):
            raise IndexError(f"control({counter}) needs samples {counter - 1} and {counter}")
        dx = self.xy[counter][0] - self.xy[counter - 1][0]
        dy = self.xy[counter][1] - self.xy[counter - 1][1]
        rot1 = math.atan2(dy, dx) - self.theta[counter - 1]
        trans = math.sqrt(dy * dy + dx * dx)
        rot2 = self.theta[counter] - self.theta[counter - 1] - rot1
        return np.array([rot1, trans, rot2])

    def controls(self) -> np.ndarray:
        """All controls, [len-1][3]."""
        return np.stack([self.control(c) for c in range(1, len(self))]) if len(self) > 1 else np.zeros((0, 3))


def read_odometry(lines: Iterable[str] | IO[str], init_pos=(0.0, 0.0)) -> OdometryTrack:
    """Read a whole odometry file.  A malformed line keeps the previous values of the fields sscanf did not reach in
    the reference; here it raises, because silently repeating a pose makes a zero-length control."""
    track = OdometryTrack(init_pos=(float(init_pos[0]), float(init_pos[1])))
    for ln, line in enumerate(lines, 1):
        if not line.strip():
            continue
        rec = parse_odometry_line(line)
        if rec is None:
            raise ValueError(f"odometry line {ln}: expected '<id> : <t> <x> <y> <theta>', got {line!r}")
        track.push(*rec)
    return track


def wrap_angle(angle: float) -> float:
    """CSLAM::wrapAngle (SLAM.cpp:507-519): one turn only, as in the reference."""
    if angle > math.pi:
        angle -= 2.0 * math.pi
    elif angle < -math.pi:
        angle += 2.0 * math.pi
    return angle


class RobotPathWriter:
    """RobotPath.txt (SLAM.cpp:3512-3562): tab-terminated fields `index, odo x, odo y, x, y, P00, P01, P10, P11`
    with "%d" / "%f" formatting; the frame with counter 1 is preceded by an all-zero pose row carrying the same
    covariance block.  P is the 2x2 block of m_P_k at the robot (x, y)."""

    def __init__(self, fp: IO[str]):
        self._fp = fp

    @staticmethod
    def _row(index: int, odo_xy, est_xy, P2) -> str:
        vals = [odo_xy[0], odo_xy[1], est_xy[0], est_xy[1], P2[0][0], P2[0][1], P2[1][0], P2[1][1]]
        return "%d\t" % index + "".join("%f\t" % float(v) for v in vals) + "\n"

    def record(self, frame_counter: int, show_counter: int, odo_xy, est_xy, P2):
        if frame_counter == 1:
            self._fp.write(self._row(1, (0.0, 0.0), (0.0, 0.0), P2))
        self._fp.write(self._row(show_counter, odo_xy, est_xy, P2))


def read_robot_path(lines: Iterable[str]) -> np.ndarray:
    """Parse RobotPath.txt back into an [rows][9] array (index first)."""
    rows = [[float(t) for t in line.split("\t") if t.strip()] for line in lines if line.strip()]
    for r in rows:
        if len(r) != 9:
            raise ValueError(f"RobotPath row with {len(r)} fields")
    return np.array(rows).reshape(-1, 9)
