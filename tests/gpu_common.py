"""Shared helpers of the GPU parity tests (all of them call the product through the C ABI only)."""
import os

import numpy as np
import pytest

DISC = [7, 12, 13, 14, 15, 16, 20, 21, 22]   # IDwp, endReached, goalReached, n_steps, fail, n_ref, tainted, trace, idwp0
CONT = [0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 17, 18, 19]
REL_TOL = 1e-6  # north-star: states and costs within 1e-6 relative; measured: 0 (the kernels evaluate glibc's libm)


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def rel_err(got, want):
    return np.abs(got - want) / np.maximum(1.0, np.abs(want))


def assert_rollouts_match(got, want, what="", exact=True):
    """Discrete outputs exactly; states and costs bit-equal (exact=True, the default: every operation is IEEE-exact or a restated
    glibc routine) or within REL_TOL."""
    d = got[:, DISC] != want[:, DISC]
    assert not d.any(), f"{what}: discrete outputs differ in rows {np.where(d.any(1))[0][:10]}"
    e = rel_err(got[:, CONT], want[:, CONT])
    assert np.nanmax(e) < REL_TOL, f"{what}: max relative error {np.nanmax(e)}"
    if exact:
        assert np.nanmax(e) == 0.0, f"{what}: states/costs not bit-equal, max relative error {np.nanmax(e)}"
    return float(np.nanmax(e))


@pytest.fixture(scope="module")
def clrrt():
    import clrrt_b200
    return clrrt_b200
