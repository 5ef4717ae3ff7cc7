"""CPU-only checks of the drop-in boundary: the shared library loads and exports every symbol include/clrrt.h
declares; struct layouts used by the Python binding match the header; host-only entry points work without a GPU."""
import ctypes as C
import os
import re

import numpy as np

import clrrt_b200 as clrrt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "clrrt.h")).read()
    return sorted(set(re.findall(r"^(?:int|const char\*)\s+(clrrt_\w+)\s*\(", hdr, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = clrrt.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libclrrt_b200.so does not export {n}"


def test_host_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "clrrt_host.h")).read()
    names = sorted(set(re.findall(r"^(?:int|void\*?)\s+(clrrt_\w+)\s*\(", hdr, flags=re.M)))
    assert len(names) >= 12
    lib = C.CDLL(os.path.join(ROOT, "cl-rrt_b200", "libclrrt_host.so"))
    for n in names:
        assert hasattr(lib, n), f"libclrrt_host.so does not export {n}"


def test_struct_layouts_match_header():
    assert clrrt.NODE_DTYPE.itemsize == 160
    assert clrrt.ROLLOUT_DTYPE.itemsize == 160
    assert clrrt.OBSTACLE_DTYPE.itemsize == 56
    assert C.sizeof(clrrt.Vehicle) == 14 * 8
    assert C.sizeof(clrrt.Params) == 14 * 8 + 11 * 8 + 5 * 8 + 4 * 8 + 8 + 3 * 8 + 8 + 8
    assert C.sizeof(clrrt.RoundStats) == 6 * 4 + 8 + 4 * 4 + 4 + 4
    assert clrrt.RECORD_BYTES == 160


def test_default_params_are_the_launch_file_and_prius():
    p = clrrt.default_params()
    assert (p.sim_dt, p.ctrl_tla, p.ctrl_mindla, p.ctrl_dlavmin, p.ctrl_Kp, p.ctrl_Ki) == (0.04, 1.4, 3.2, 3, 8, 0.05)
    assert list(p.Wcost) == [10, 5, 0, 4, 1] and p.ref_res == 0.2 and p.vmax == 5
    assert p.veh.L == 2.7 and p.veh.dmax == 0.52 and abs(p.veh.Kus - 0.0139629346) < 1e-9


def test_sampling_matches_oracle_and_known_answer():
    from cpulib import CpuPlanner
    s, h = clrrt.draw_samples((50, 0, 0, 0), 64, seed=1)
    o = CpuPlanner("oracle")
    o.srand(1)
    o.tree_init()
    s2, h2, _ = o.draw_samples(64)
    assert np.array_equal(s, s2) and np.array_equal(h, h2)
    assert abs(s[0, 0] - 50.41126251220703) < 1e-12 and h[0] == 1


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        clrrt.Planner()
    except clrrt.ClrrtError as e:
        assert "clrrt_create failed" in str(e)
    else:
        raise AssertionError("Planner() succeeded without a GPU")
