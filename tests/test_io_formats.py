"""odometry reader / control conversion / RobotPath writer (SLAM.cpp:462-496, 1444-1454, 3512-3562)."""
import io
import math

import numpy as np
import pytest

from cv_monoslam_b200 import io_formats as F


def test_parse_line_matches_sscanf_format():
    assert F.parse_odometry_line("12 : 1354.25 0.5 -1.25 0.125\n") == (12, 0.5, -1.25, 0.125)
    assert F.parse_odometry_line("  7:3 1e-2 2E+1 -.5") == (7, 0.01, 20.0, -0.5)
    assert F.parse_odometry_line("7 : 3 1 2") is None
    assert F.parse_odometry_line("x : 3 1 2 3") is None


def test_track_rebases_on_initial_position():
    text = "0 : 0.0 10.0 20.0 0.0\n1 : 0.1 11.0 20.0 0.0\n2 : 0.2 11.0 22.0 1.5\n"
    tr = F.read_odometry(io.StringIO(text), init_pos=(-1.0, 2.0))
    assert len(tr) == 3 and tr.ids == [0, 1, 2]
    assert tr.xy == [(-1.0, 2.0), (0.0, 2.0), (0.0, 4.0)]
    assert tr.theta == [0.0, 0.0, 1.5]


def test_control_matches_oracle_odometry_to_control(oracle):
    rng = np.random.default_rng(0)
    poses = np.cumsum(rng.normal(0, [0.05, 0.05, 0.02], (20, 3)), axis=0)
    tr = F.OdometryTrack(init_pos=(0.3, -0.2))
    for i, (x, y, t) in enumerate(poses):
        tr.push(i, x, y, t)
    U = tr.controls()
    assert U.shape == (19, 3)
    for c in range(1, 20):
        prev = np.array([tr.xy[c - 1][0], tr.xy[c - 1][1], tr.theta[c - 1]])
        now = np.array([tr.xy[c][0], tr.xy[c][1], tr.theta[c]])
        assert np.array_equal(U[c - 1], oracle.odometry_to_control(prev, now))
    with pytest.raises(IndexError):
        tr.control(0)
    with pytest.raises(IndexError):
        tr.control(20)


def test_malformed_line_raises():
    with pytest.raises(ValueError):
        F.read_odometry(["0 : 0 1 2 3\n", "garbage\n"])


def test_wrap_angle_single_turn():
    assert F.wrap_angle(math.pi + 0.5) == pytest.approx(-math.pi + 0.5)
    assert F.wrap_angle(-math.pi - 0.5) == pytest.approx(math.pi - 0.5)
    assert F.wrap_angle(0.3) == 0.3
    assert F.wrap_angle(3 * math.pi + 0.1) == pytest.approx(math.pi + 0.1)   # one turn only, like the reference


def test_robot_path_layout_and_round_trip():
    buf = io.StringIO()
    w = F.RobotPathWriter(buf)
    P = [[0.25, -0.5], [-0.5, 4.0]]
    w.record(1, 1, (1.5, 2.5), (1.25, 2.75), P)
    w.record(2, 2, (1.75, 2.5), (1.5, 2.875), P)
    lines = buf.getvalue().splitlines(keepends=True)
    assert lines[0] == "1\t0.000000\t0.000000\t0.000000\t0.000000\t0.250000\t-0.500000\t-0.500000\t4.000000\t\n"
    assert lines[1] == "1\t1.500000\t2.500000\t1.250000\t2.750000\t0.250000\t-0.500000\t-0.500000\t4.000000\t\n"
    assert len(lines) == 3
    arr = F.read_robot_path(lines)
    assert arr.shape == (3, 9)
    assert np.array_equal(arr[2], [2, 1.75, 2.5, 1.5, 2.875, 0.25, -0.5, -0.5, 4.0])
