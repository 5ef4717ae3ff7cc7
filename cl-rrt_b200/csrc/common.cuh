// common.cuh — device-side types shared by the CL-RRT kernels (sm_100a).
//
// Data layout in HBM: the tree is a structure of arrays, one array per node field, so that
// the nearest-node kernel streams only the fields it needs with coalesced loads and the rollout kernel
// gathers a parent's ~15 scalars once per rollout.  The same SoA type is used for the per-round staging
// area that holds freshly accepted nodes before they are appended in sample order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/clrrt.h"

#define CLRRT_WARP 32
#define FULL_MASK 0xffffffffu

// ---- std::min / std::max exactly as libstdc++ defines them (NaN behaviour differs from fmin/fmax) -------
// std::max(a,b) = (a<b)?b:a ; std::min(a,b) = (b<a)?b:a.  The reference relies on max(0.0, NaN) == 0
// (rrt/src/reference.cpp:146) and lets NaN pass through its saturation helpers.
template <typename T> __device__ __forceinline__ T std_max(T a, T b) { return (a < b) ? b : a; }
template <typename T> __device__ __forceinline__ T std_min(T a, T b) { return (b < a) ? b : a; }
// checkSaturation(min,max,val) rrt/include/rrt/functions.h:60-62 == enforceConstraints rrt/src/simulation.cpp:7-9
template <typename T> __device__ __forceinline__ T saturate(T lo, T hi, T val) { return std_max(std_min(val, hi), lo); }
template <typename T> __device__ __forceinline__ T sq(T x) { return x * x; }  // pow(x,2)

// ---- node fields, structure of arrays ------------------------------------------------------------------
struct NodeSoA {
  double *x, *y, *th, *de, *v, *a, *t;  // state[0..6]
  double *s7, *s8, *s9;                 // state[7..9] (logging slots: IDwp, v_ref, delta_cmd)
  double *rfx, *rfy, *rbx, *rby;        // ref.{x,y}.front(), ref.{x,y}.back()
  double *vback;                        // ref.v.back()
  double *smx, *smy;                    // the sample this node's reference was aimed at (re-materialisation)
  float *costE, *costS;                 // float, as in struct Node
  int32_t *parent, *goal, *nref, *kind;
  // derived, for the nearest-node kernel (filled by derive_nodes_kernel)
  float *ca, *sa;                       // cosf/sinf of ang = (float)(-theta)   rrt/src/rrtplanner.cpp:378-380
  double *angPar;                       // atan2(ref back - ref front)          rrt/src/rrtplanner.cpp:273
};
#define NODE_SOA_DOUBLE_FIELDS 18
#define NODE_SOA_FLOAT_FIELDS 4
#define NODE_SOA_INT_FIELDS 4

// ---- obstacle tables -------------------------------------------------------------------------------------
// Static obstacles (vel == 0): the OBB of rrt/src/old_collisioncheck.cpp:6-22 does not depend on time, so the
// float vertices (setVertices :56-65), axes (setNorms :67-76) and the obstacle's own projection intervals are
// computed ONCE on the host, with the host libm's cosf/sinf (bit-identical to what the reference computes).
struct __align__(16) ObsHot {   // read for every (lane, obstacle) pair: 32 B
  float vx[4], vy[4];
};
struct __align__(16) ObsCold {  // read only when the vehicle's own axes did not separate the pair: 64 B
  float nx[4], ny[4];           // axes; ny[3] = 0 (never written upstream, "defined" variant)
  float pmax[4], pmin[4];       // the obstacle's own projection interval on each of its axes
};
// Broad-phase record (verdict-only mode, see rollout.cuh): bounding circle of a static obstacle.  The centre is
// stored RELATIVE to the grid origin (subtracted in double on the host) so that float keeps ~1e-4 m over any scene.
struct __align__(16) ObsBound {
  float cx, cy, rr, ohh;   // rr = half diagonal of the obstacle box + margin; ohh/ohw = half extents along / across
  float oc, os, ohw, pad;  // the obstacle's heading (oc, os) = (cosf, sinf) of it
};
// Moving obstacles: centre = c + vel*t with t = x[6] of the lane, so vertices are rebuilt per step from
// host-computed float half-extent products (the float operation order of setVertices is preserved).
struct __align__(16) ObsMoving {
  double cx, cy, vx, vy;
  float ch, sw, sh, cw;  // cosf(o)*(h/2), sinf(o)*(w/2), sinf(o)*(h/2), cosf(o)*(w/2)
  float rr, ohh, pad0, pad1;  // broad phase: rr = half diagonal + margin (+ slack for the float prediction), half length
  float oc, os, ohw, pad2;    //              heading (cosf, sinf), half width — same layout as ObsBound's second half
};

// ---- parameters in constant memory -----------------------------------------------------------------------
struct DevParams {
  // Vehicle (rrt/include/rrt/vehicle.h) — only what the rollout reads
  double dmax, ddmax, inv_Td, inv_Ta, amin, amax, L, Vch, Kus;
  // controller / simulation globals
  double sim_dt, mindla, tla, dla_c, Kp, Ki, ref_res, vmax, ay_road_max;
  double W[5];
  double goal[4];
  double feas_len;      // 2.1*ref_res, rrt/src/rrtplanner.cpp:283
  // goal-bias geometry, rrt/src/rrtplanner.cpp:292-315 and rrt/src/reference.cpp:25-54 (host-evaluated with libm)
  double gb_clx, gb_cly, gb_crx, gb_cry, gb_R2;
  double gb_P1x, gb_P1y, gb_P2x, gb_P2y, gb_ext_x, gb_ext_y;
  // vehicle box of rrt/src/old_collisioncheck.cpp:34-36
  float veh_hw, veh_hh;  // w/2, h/2 with w=2.0f, h=(float)4.848
  int32_t max_steps;     // number of i with i < 20/sim_dt, rrt/src/simulation.cpp:58
  int32_t obs_use_pred;
  int32_t n_static, n_moving;
  int32_t static_in_smem;
  // uniform grid over vehicle-box-centre positions (broad phase, verdict-only mode): cell -> list of the static
  // obstacles whose bounding circle can come within reach of a vehicle centred anywhere in the cell
  int32_t grid_nx, grid_ny;
  float grid_inv_cell;
  double grid_ox, grid_oy;
  // pose grid (x, y, heading mod pi), pose_sub x pose_sub cells per position-grid cell; pose_nh == 0: not built
  int32_t pose_sub, pose_nh;
  int32_t exact_dist;    // 1: return the reference's pseudo-distance (needed when W[2] != 0); 0: verdict only
  int32_t bend;          // curved-road mode: lane-deviation cost per step (rrt/src/simulation.cpp:92-95)
  double lane_S, Cxy1, Cxy2;
  float veh_reach;       // half diagonal of the vehicle box (broad phase)
  // verdict-only collision check (rollout.cuh, warp_collide): a gap above fine_margin along one of the four box directions
  // is a gap the reference's float SAT sees; an overlap above deep_margin along all four is an intersection it sees; only
  // pairs in between run the reference's SAT.  Both scale with the scene's coordinates (clrrt_set_obstacles).
  float fine_margin, deep_margin;
};

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
