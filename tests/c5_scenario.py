"""Config C5: a 5 Hz receding-horizon loop.  The world is fixed; every query hands the planner its inputs
in the CAR frame, as the reference's mission planner and detection node do: goal, obstacles (centre, heading, velocity
rotated), car state [0, 0, 0, delta, v, a].  Between queries the car is advanced 0.2 s along the previous best path and
moving obstacles by their velocity.  TEST INFRASTRUCTURE (shared by tests/golden/make_golden.py and the GPU test)."""
import numpy as np

def world_goal(q):
    """A goal that recedes along the road, as a mission planner would hand out: ~55-60 m ahead of the car."""
    return np.array([60.0 + 0.9 * q, 0.0, 0.0, 0.0])


DT_QUERY = 0.2
SIM_DT = 0.04


def world_obstacles(t):
    """C1 boxes, odd ones driving towards the car at 1 m/s."""
    o = np.zeros((10, 7))
    for i in range(10):
        vx = -1.0 if i % 2 == 1 else 0.0
        o[i] = [8 + 4.7 * i + vx * t, 3.0 if i % 2 == 0 else -3.0, 0.0, 4.0, 8.0, vx, 0.0]
    return o


def to_car_frame(world_state, goal_w, obs_w):
    x, y, th = world_state[:3]
    c, s = np.cos(th), np.sin(th)

    def pt(px, py):
        return (px - x) * c + (py - y) * s, -(px - x) * s + (py - y) * c

    gx, gy = pt(goal_w[0], goal_w[1])
    goal = np.array([gx, gy, goal_w[2] - th, goal_w[3]])
    obs = obs_w.copy()
    for k in range(len(obs)):
        obs[k, 0], obs[k, 1] = pt(obs_w[k, 0], obs_w[k, 1])
        obs[k, 2] = obs_w[k, 2] - th
        obs[k, 5] = obs_w[k, 5] * c + obs_w[k, 6] * s
        obs[k, 6] = -obs_w[k, 5] * s + obs_w[k, 6] * c
    return goal, obs


def advance(world_state, best_traj_world, rows_per_node):
    """The state 0.2 s further down the best path: the trajectory row closest to the car, plus five sim steps.
    Trajectory rows hold world x, y; their headings are left in whatever frame they were simulated in (the reference
    does not rotate them, rrt/src/transformations.cpp:296-299, :310-313), so the harness takes the heading from the
    path tangent instead."""
    steps = int(round(DT_QUERY / SIM_DT))
    n = int(np.sum(rows_per_node))
    rows = np.asarray(best_traj_world[:n], float)
    if len(rows) < 3:
        return np.array(world_state, float)
    d2 = (rows[:, 0] - world_state[0]) ** 2 + (rows[:, 1] - world_state[1]) ** 2
    i = min(int(np.argmin(d2)) + steps, len(rows) - 2)
    i = max(i, 1)
    r = rows[i]
    th = np.arctan2(rows[i + 1, 1] - rows[i - 1, 1], rows[i + 1, 0] - rows[i - 1, 0])
    return np.array([r[0], r[1], th, r[3], r[4], r[5]])
