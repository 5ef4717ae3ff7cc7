"""Collision verdicts at the boundary.  Every rollout step of the product takes the verdict-only check: pose grid, a
three-way classification along the four box directions (separated by more than fine_margin / overlapping by more than
deep_margin / in between) and the reference's float SAT only for the band in between (rollout.cuh, box_class).  The
margin argument is exercised here on purpose: poses are placed within +-1 cm of TOUCHING an obstacle (found by bisection
on an exact double-precision overlap test), many of them within micrometres, and clrrt_collide_batch's verdict is compared
with checkObsDistance == 0 of the oracle (and of the reference binary when present), pose by pose: zero disagreements."""
import numpy as np
import pytest

from cpulib import CpuPlanner, ref_available, scene_c1_boxes, scene_c3_boxes
from gpu_common import clrrt  # noqa: F401

pytestmark = pytest.mark.gpu
VEH_HL, VEH_HW, VEH_OFF = 4.848 / 2, 1.0, 1.424  # vehicle box of rrt/src/old_collisioncheck.cpp:34-36


def rect_axes(th):
    c, s = np.cos(th), np.sin(th)
    return np.stack([c, s], -1), np.stack([-s, c], -1)


def overlap(vc, vth, oc, oth, ohl, ohw):
    """Exact-arithmetic-ish (double) SAT of two rectangles, vectorised: True where they intersect."""
    d = oc - vc
    va, vb = rect_axes(vth)
    oa, ob = rect_axes(oth)
    sep = np.zeros(len(vc), bool)
    for ax in (va, vb, oa, ob):
        rv = VEH_HL * np.abs((va * ax).sum(-1)) + VEH_HW * np.abs((vb * ax).sum(-1))
        ro = ohl * np.abs((oa * ax).sum(-1)) + ohw * np.abs((ob * ax).sum(-1))
        sep |= np.abs((d * ax).sum(-1)) > rv + ro
    return ~sep


def near_touching_poses(obs, n, rng, t_max=0.0, band=0.01, corridor=False):
    """n rear-axle poses (x, y, theta, t) whose vehicle box is within +-band metres (along a random approach direction) of
    touching a randomly chosen obstacle of `obs`; a third of them within 1e-6 m, a few exactly at the bisection point."""
    k = rng.integers(0, 200 if corridor else len(obs), n)  # corridor: the two rows that border the free lane of C3
    o = obs[k]
    t = rng.uniform(0, t_max, n) if t_max > 0 else np.zeros(n)
    oc = o[:, 0:2] + o[:, 5:7] * t[:, None]
    oth = o[:, 2]
    ohl, ohw = o[:, 4] / 4, o[:, 3] / 4  # effective box: size_y/2 long (along theta) x size_x/2 wide, SURVEY 8a row 23
    vth = rng.uniform(-np.pi, np.pi, n)
    ua = rng.uniform(-np.pi, np.pi, n)
    if corridor:  # approach from inside the lane with a heading roughly along it, as rollouts do
        vth = rng.uniform(-0.5, 0.5, n) + np.where(rng.random(n) < 0.5, 0.0, np.pi)
        ua = -np.sign(o[:, 1]) * rng.uniform(0.6, np.pi - 0.6, n)
    u = np.stack([np.cos(ua), np.sin(ua)], -1)
    lo, hi = np.zeros(n), np.full(n, 12.0)  # centre distance along u: overlapping at 0, separated at 12 m
    for _ in range(48):
        mid = 0.5 * (lo + hi)
        ov = overlap(oc + u * mid[:, None], vth, oc, oth, ohl, ohw)
        lo = np.where(ov, mid, lo)
        hi = np.where(ov, hi, mid)
    kind = rng.integers(0, 6, n)
    off = np.where(kind < 2, rng.uniform(-1e-6, 1e-6, n), np.where(kind < 3, rng.uniform(-1e-4, 1e-4, n), rng.uniform(-band, band, n)))
    off = np.where(rng.random(n) < 0.02, 0.0, off)
    vc = oc + u * (lo + off)[:, None]
    pos = vc - VEH_OFF * np.stack([np.cos(vth), np.sin(vth)], -1)
    return np.column_stack([pos, vth, t])


def check(clrrt, obs, poses, what, min_each=0.15):
    orc = CpuPlanner("oracle")
    orc.set_obstacles(obs)
    want = orc.obs_distance_batch(poses)
    pl = clrrt.Planner(device=0, tree_capacity=64, max_round=64)
    pl.set_obstacles(obs)
    got, dist = pl.collide_batch(poses, want_distance=True)
    # after a parameter update too (the margins must survive clrrt_set_params)
    pl.set_query((0, 0, 0, 0, 1, 0), (60, 1, 0.1, 0), 4.0)
    got2 = pl.collide_batch(poses)
    pl.close()
    hit = want == 0
    frac = hit.mean()
    assert min_each < frac < 1 - min_each, f"{what}: {frac:.2f} of the poses collide — the fuzz is not at the boundary"
    bad = np.where(got.astype(bool) != hit)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} verdicts differ from checkObsDistance == 0, first poses {poses[bad[:3]]}"
    assert np.array_equal(got, got2), f"{what}: verdicts changed after clrrt_set_params"
    # the exact path returns the reference's value itself
    assert np.array_equal(dist, want), f"{what}: pseudo-distance differs in {(dist != want).sum()} poses"
    if ref_available(True):
        ref = CpuPlanner("ref_defined")
        ref.set_obstacles(obs)
        sub = slice(0, min(len(poses), 20000))
        assert np.array_equal(ref.obs_distance_batch(poses[sub]), want[sub]), f"{what}: oracle and reference binary disagree"
    return int(hit.sum())


def test_dense_scene_boundary(clrrt):
    """C3: 1000 boxes; 200 000 poses within 1 cm of touching one of them."""
    obs = scene_c3_boxes()
    rng = np.random.default_rng(11)
    poses = np.concatenate([near_touching_poses(obs, 150_000, rng, corridor=True), near_touching_poses(obs, 50_000, rng)])
    check(clrrt, obs, poses, "C3 dense scene", min_each=0.03)  # neighbours collide too in the dense field


def test_static_boxes_boundary(clrrt):
    obs = scene_c1_boxes()
    poses = near_touching_poses(obs, 400_000, np.random.default_rng(12))
    check(clrrt, obs, poses, "C1 boxes")


def test_moving_boxes_boundary(clrrt):
    obs = scene_c1_boxes(moving=True)
    poses = near_touching_poses(obs, 300_000, np.random.default_rng(13), t_max=20.0)
    check(clrrt, obs, poses, "C1 boxes, odd ones moving")


def test_far_origin_boundary(clrrt):
    """The same boxes 10 km from the origin: float vertices carry ~1 mm of rounding there and the margins scale with it."""
    obs = scene_c1_boxes()
    obs[:, 0] += 1.0e4
    obs[:, 1] -= 1.0e4
    rng = np.random.default_rng(14)
    obs[:, 2] = rng.uniform(0, np.pi, len(obs))
    poses = near_touching_poses(obs, 100_000, rng, band=0.05)
    check(clrrt, obs, poses, "10 km from the origin")


def test_random_boxes_all_orientations(clrrt):
    rng = np.random.default_rng(15)
    obs = np.zeros((200, 7))
    obs[:, 0] = rng.uniform(-40, 80, 200); obs[:, 1] = rng.uniform(-40, 40, 200); obs[:, 2] = rng.uniform(-np.pi, np.pi, 200)
    obs[:, 3] = rng.uniform(0.4, 6, 200); obs[:, 4] = rng.uniform(0.4, 10, 200)
    mv = rng.random(200) < 0.3
    obs[mv, 5] = rng.uniform(-2, 2, mv.sum()); obs[mv, 6] = rng.uniform(-1, 1, mv.sum())
    poses = near_touching_poses(obs, 200_000, rng, t_max=15.0)
    check(clrrt, obs, poses, "200 random boxes", min_each=0.05)
