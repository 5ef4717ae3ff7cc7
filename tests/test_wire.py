"""ROS 1 wire format of the planner's messages (cl-rrt_b200/host/clrrt_wire.cpp): byte-exact against the
layout the .msg files define (car_msgs/msg/*.msg, car_msgs/srv/getobstacles.srv), built here with `struct`.  CPU only."""
import ctypes as C
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lib():
    l = C.CDLL(os.path.join(ROOT, "cl-rrt_b200", "libclrrt_host.so"))
    return l


def f64s(v):
    return struct.pack("<I", len(v)) + struct.pack(f"<{len(v)}d", *v)


def test_motion_request_and_state_parse():
    l = lib()
    goal, vmax, bend, cxy, cxs, shifts = [50.0, 1.5, 0.1, 2.0], 5.0, False, [0.0, 0.01, 0.2], [1.0, 0.0], [0.0, 3.5]
    buf = f64s(goal) + struct.pack("<d", vmax) + struct.pack("<B", int(bend)) + f64s(cxy) + f64s(cxs) + f64s(shifts)
    g = (C.c_double * 4)()
    v, b, n = C.c_double(0), C.c_int(-1), C.c_int(-1)
    assert l.clrrt_wire_parse_request(buf, len(buf), g, C.byref(v), C.byref(b), C.byref(n)) == 0
    assert list(g) == goal and v.value == vmax and b.value == 0 and n.value == 2
    # truncated buffers and short goals are refused
    assert l.clrrt_wire_parse_request(buf[:-3], len(buf) - 3, g, C.byref(v), C.byref(b), C.byref(n)) != 0
    bad = f64s(goal[:3]) + struct.pack("<d", vmax) + b"\x00" + f64s([]) + f64s([]) + f64s([])
    assert l.clrrt_wire_parse_request(bad, len(bad), g, C.byref(v), C.byref(b), C.byref(n)) != 0
    state = [1.0, -2.0, 0.3, 0.01, 4.5, -0.2]
    s6 = (C.c_double * 6)()
    sb = f64s(state)
    assert l.clrrt_wire_parse_state(sb, len(sb), s6) == 0 and list(s6) == state
    assert l.clrrt_wire_parse_state(f64s(state[:5]), len(f64s(state[:5])), s6) != 0


def test_trajectory_bytes():
    l = lib()
    rows = np.arange(24, dtype=np.float64).reshape(3, 8) * 0.25
    out = (C.c_uint8 * 4096)()
    n = l.clrrt_wire_trajectory(rows.ctypes.data_as(C.c_void_p), 3, out, 4096)
    want = b"".join(f64s(rows[:, k].tolist()) for k in range(8))  # x y theta delta v a a_cmd d_cmd
    assert n == len(want) and bytes(out[:n]) == want
    assert l.clrrt_wire_trajectory(rows.ctypes.data_as(C.c_void_p), 3, out, 10) < 0


def test_obstacles_roundtrip():
    l = lib()
    obs = np.array([[8.0, 3.0, 0.1, 4.0, 8.0, -1.0, 0.0], [12.7, -3.0, 0.0, 2.0, 4.0, 0.0, 0.5]])
    out = (C.c_uint8 * 4096)()
    n = l.clrrt_wire_obstacles(obs.ctypes.data_as(C.c_void_p), 2, out, 4096)
    # Obstacle2D = Pose2D centre (x, y, theta), size_x, size_y, Twist (linear xyz, angular xyz): 11 doubles
    want = struct.pack("<I", 2)
    for o in obs:
        want += struct.pack("<11d", o[0], o[1], o[2], o[3], o[4], o[5], o[6], 0.0, 0.0, 0.0, 0.0)
    assert n == len(want) and bytes(out[:n]) == want
    back = np.zeros((4, 7))
    assert l.clrrt_wire_parse_obstacles(bytes(out[:n]), n, back.ctypes.data_as(C.c_void_p), 4) == 2
    assert np.array_equal(back[:2], obs)
    assert l.clrrt_wire_parse_obstacles(bytes(out[:n - 1]), n - 1, back.ctypes.data_as(C.c_void_p), 4) < 0


def test_motion_response_layout():
    """preparePathMessage (rrt/src/motionplanner.cpp:235-261): per segment one Reference (all points) and one
    Trajectory that skips the segment's first state; serialised as Reference[] then Trajectory[]."""
    l = lib()
    rows = np.arange(50, dtype=np.float64).reshape(5, 10)
    per = np.array([2, 3], np.int32)
    out = (C.c_uint8 * 8192)()
    n = l.clrrt_wire_response_roundtrip(rows.ctypes.data_as(C.c_void_p), per.ctypes.data_as(C.c_void_p), 2, out, 8192)
    assert n > 0
    segs = [rows[:2], rows[2:]]
    want = struct.pack("<I", 2)
    for s in segs:
        want += f64s(s[:, 0].tolist()) + f64s(s[:, 1].tolist()) + f64s(s[:, 4].tolist()) + struct.pack("<i", 1)
    want += struct.pack("<I", 2)
    for s in segs:
        t = s[1:]
        for k in (0, 1, 2, 3, 4, 5, 8, 9):
            want += f64s(t[:, k].tolist())
    assert bytes(out[:n]) == want
