"""Curved-road mode on the host: sampleOnLane and the road-frame transforms of
rrt/src/transformations.cpp:20-202, through libclrrt_host.so.  Against the reference's own functions (oracle/_ref) when that
build is present: bit-equal.  Always: the round-trip and geometric properties the transforms must have, and golden values
generated from the reference (tests/golden/g7_road.npz, tests/golden/make_golden.py g7)."""
import ctypes as C
import os

import numpy as np
import pytest

from cpulib import ref_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CXY = np.array([0.004, 0.05, 1.5])     # road y = 0.004 x^2 + 0.05 x + 1.5 (radius ~125 m)
LANES = np.array([-3.5, 0.0, 3.5])


def arc_length_fit(Cxy, xmax=80.0):
    """Cxs: second-order fit of the road's arc length S(x), as the reference's road model supplies it."""
    x = np.linspace(0, xmax, 400)
    dy = 2 * Cxy[0] * x + Cxy[1]
    S = np.concatenate([[0], np.cumsum(np.sqrt(1 + ((dy[1:] + dy[:-1]) / 2) ** 2) * np.diff(x))])
    return np.polyfit(x, S, 2)


CXS = arc_length_fit(CXY)


def host():
    lib = C.CDLL(os.path.join(ROOT, "cl-rrt_b200", "libclrrt_host.so"))
    lib.clrrt_host_road_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.clrrt_host_sample_on_lane.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def transform(lib, fn, what, recs):
    r = np.ascontiguousarray(recs, dtype=np.float64).copy()
    assert getattr(lib, fn)(what, CXY.ctypes.data, CXS.ctypes.data, r.ctypes.data, len(r)) == 0
    return r


def cases(n=4000, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.uniform(2, 70, n)
    y = CXY[0] * x ** 2 + CXY[1] * x + CXY[2] + rng.uniform(-6, 6, n)
    return np.column_stack([x, y, rng.uniform(-0.6, 0.6, n), rng.uniform(-0.3, 0.3, n)])


def test_round_trips_and_geometry():
    lib = host()
    p = cases()
    # the foot point lies on the parabola and the connecting segment is normal to the road there
    a = transform(lib, "clrrt_host_road_transform", 6, p)
    assert np.allclose(a[:, 1], CXY[0] * a[:, 0] ** 2 + CXY[1] * a[:, 0] + CXY[2], atol=1e-9)
    tang = np.column_stack([np.ones(len(a)), 2 * CXY[0] * a[:, 0] + CXY[1]])
    dot = ((p[:, :2] - a[:, :2]) * tang).sum(1) / np.linalg.norm(tang, axis=1)
    assert np.abs(dot).max() < 2e-3   # the closed form runs in float upstream
    # car -> road -> car returns the point (to the accuracy of the float arc projection and of the arc-length fit)
    r = transform(lib, "clrrt_host_road_transform", 0, p)
    back = transform(lib, "clrrt_host_road_transform", 1, r)
    assert np.abs(back[:, :2] - p[:, :2]).max() < 0.05
    # the lateral coordinate survives: distance to the straightened road == distance to the arc, same side
    th = np.arctan2(CXY[1], 1)
    rho = -np.sin(th) * r[:, 0] + np.cos(th) * (r[:, 1] - CXY[2])
    d_arc = np.linalg.norm(p[:, :2] - a[:, :2], axis=1) * np.sign(p[:, 1] - (CXY[0] * p[:, 0] ** 2 + CXY[1] * p[:, 0] + CXY[2]))
    assert np.abs(rho - d_arc).max() < 1e-6
    # poses: heading relative to the road is kept; states: the road-curvature steer angle is subtracted and added back
    q = transform(lib, "clrrt_host_road_transform", 2, p)
    assert ((q[:, 2] >= 0) & (q[:, 2] <= 2 * np.pi)).all()
    rel_in = p[:, 2] - np.arctan2(2 * CXY[0] * a[:, 0] + CXY[1], 1)
    rel_out = q[:, 2] - th
    assert np.abs(np.angle(np.exp(1j * (rel_in - rel_out)))).max() < 1e-9
    s = transform(lib, "clrrt_host_road_transform", 4, p)
    s2 = transform(lib, "clrrt_host_road_transform", 5, s)
    assert np.abs(s2[:, 3] - p[:, 3]).max() < 1e-4 and np.abs(s2[:, :2] - p[:, :2]).max() < 0.05


def test_sample_on_lane_is_on_a_lane_centre_line():
    lib = host()
    K = 2000
    s, h = np.zeros((K, 2)), np.zeros(K, np.uint8)
    C.CDLL(None).srand(C.c_uint(1))
    assert lib.clrrt_host_sample_on_lane(CXY.ctypes.data, LANES.ctypes.data, len(LANES), 60.0, 4.0, K, s.ctypes.data, h.ctypes.data) == 0
    th = np.arctan2(CXY[1], 1)
    S = np.cos(th) * s[:, 0] + np.sin(th) * (s[:, 1] - CXY[2])
    rho = -np.sin(th) * s[:, 0] + np.cos(th) * (s[:, 1] - CXY[2])
    dla = max(3.2, 3.2 - 1.4 * 3 + 1.4 * 4.0)
    assert S.min() >= dla - 1e-9 and S.max() <= 60.0 + 1e-4
    assert np.abs(rho[:, None] - LANES[None, :]).min(1).max() < 1e-9
    assert set(np.round(rho, 6)) == set(LANES) and 0.2 < h.mean() < 0.4   # every lane is drawn; 30 % optimise draws


def test_golden_values(golden_dir):
    """Outputs of the reference's own functions on fixed inputs (generated here from oracle/_ref)."""
    g = np.load(os.path.join(golden_dir, "g7_road.npz"))
    lib = host()
    global CXY, CXS
    cxy, cxs = CXY, CXS
    try:
        CXY, CXS = np.ascontiguousarray(g["Cxy"]), np.ascontiguousarray(g["Cxs"])
        for what in range(7):
            got = transform(lib, "clrrt_host_road_transform", what, g["inputs"])
            assert np.array_equal(got, g[f"out{what}"]), f"transform {what} differs from the reference's golden output"
        K = len(g["lane_samples"])
        s, h = np.zeros((K, 2)), np.zeros(K, np.uint8)
        C.CDLL(None).srand(C.c_uint(7))
        lanes = np.ascontiguousarray(g["lanes"])
        lib.clrrt_host_sample_on_lane(CXY.ctypes.data, lanes.ctypes.data, len(lanes), 60.0, 4.0, K, s.ctypes.data, h.ctypes.data)
        assert np.array_equal(s, g["lane_samples"]) and np.array_equal(h, g["lane_heuristic"])
    finally:
        CXY, CXS = cxy, cxs


@pytest.mark.skipif(not ref_available(False), reason="oracle/_ref not built")
def test_against_the_reference_binary():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libclrrt_ref.so"))
    ref.ref_init()
    ref.ref_road_transform.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    ref.ref_sample_on_lane.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    lib = host()
    p = cases(3000, seed=9)
    for what in range(7):
        src = p if what in (0, 2, 4, 6) else transform(lib, "clrrt_host_road_transform", what - 1, p)
        assert np.array_equal(transform(lib, "clrrt_host_road_transform", what, src), transform(ref, "ref_road_transform", what, src)), what
    K = 500
    out = []
    for f, l in ((lib.clrrt_host_sample_on_lane, lib), (ref.ref_sample_on_lane, ref)):
        s, h = np.zeros((K, 2)), np.zeros(K, np.uint8)
        C.CDLL(None).srand(C.c_uint(5))
        f(CXY.ctypes.data, LANES.ctypes.data, len(LANES), 60.0, 4.0, K, s.ctypes.data, h.ctypes.data)
        out.append((s, h))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
