// clrrt_wire.cpp — see clrrt_wire.hpp.
#include "clrrt_wire.hpp"

#include <cstring>

namespace clrrt {
namespace wire {
namespace {
const uint32_t kMaxArray = 1u << 24;  // refuse absurd counts in corrupt buffers

struct Writer {
  Bytes b;
  void raw(const void* p, size_t n) { const uint8_t* q = static_cast<const uint8_t*>(p); b.insert(b.end(), q, q + n); }
  void u32(uint32_t v) { raw(&v, 4); }     // the hosts this runs on are little endian, like the ROS wire
  void i32(int32_t v) { raw(&v, 4); }
  void f64(double v) { raw(&v, 8); }
  void u8(uint8_t v) { raw(&v, 1); }
  void f64s(const std::vector<double>& v) { u32((uint32_t)v.size()); if (!v.empty()) raw(v.data(), 8 * v.size()); }
};
struct Reader {
  const uint8_t* p;
  size_t n, off = 0;
  bool ok = true;
  bool need(size_t k) { if (!ok || n - off < k) { ok = false; return false; } return true; }
  uint32_t u32() { uint32_t v = 0; if (need(4)) { memcpy(&v, p + off, 4); off += 4; } return v; }
  int32_t i32() { int32_t v = 0; if (need(4)) { memcpy(&v, p + off, 4); off += 4; } return v; }
  double f64() { double v = 0; if (need(8)) { memcpy(&v, p + off, 8); off += 8; } return v; }
  uint8_t u8() { uint8_t v = 0; if (need(1)) { v = p[off]; off += 1; } return v; }
  void f64s(std::vector<double>& v) {
    const uint32_t k = u32();
    v.clear();
    if (!ok || k > kMaxArray || !need(8 * (size_t)k)) { ok = false; return; }
    v.resize(k);
    if (k) memcpy(v.data(), p + off, 8 * (size_t)k);
    off += 8 * (size_t)k;
  }
};
void put(Writer& w, const MotionRequest& m) {
  w.f64s(m.goal); w.f64(m.vmax); w.u8(m.bend ? 1 : 0); w.f64s(m.Cxy); w.f64s(m.Cxs); w.f64s(m.laneShifts);
}
void get(Reader& r, MotionRequest& m) {
  r.f64s(m.goal); m.vmax = r.f64(); m.bend = r.u8() != 0; r.f64s(m.Cxy); r.f64s(m.Cxs); r.f64s(m.laneShifts);
}
void put(Writer& w, const Trajectory& m) {
  w.f64s(m.x); w.f64s(m.y); w.f64s(m.theta); w.f64s(m.delta); w.f64s(m.v); w.f64s(m.a); w.f64s(m.a_cmd); w.f64s(m.d_cmd);
}
void get(Reader& r, Trajectory& m) {
  r.f64s(m.x); r.f64s(m.y); r.f64s(m.theta); r.f64s(m.delta); r.f64s(m.v); r.f64s(m.a); r.f64s(m.a_cmd); r.f64s(m.d_cmd);
}
void put(Writer& w, const MyReference& m) { w.f64s(m.x); w.f64s(m.y); w.f64s(m.v); w.i32(m.dir); }
void get(Reader& r, MyReference& m) { r.f64s(m.x); r.f64s(m.y); r.f64s(m.v); m.dir = r.i32(); }
void put(Writer& w, const Obstacle2D& o) {
  w.f64(o.obb.center.x); w.f64(o.obb.center.y); w.f64(o.obb.center.theta); w.f64(o.obb.size_x); w.f64(o.obb.size_y);
  w.f64(o.vel.linear.x); w.f64(o.vel.linear.y); w.f64(o.vel.linear.z);
  w.f64(o.vel.angular.x); w.f64(o.vel.angular.y); w.f64(o.vel.angular.z);
}
void get(Reader& r, Obstacle2D& o) {
  o.obb.center.x = r.f64(); o.obb.center.y = r.f64(); o.obb.center.theta = r.f64(); o.obb.size_x = r.f64(); o.obb.size_y = r.f64();
  o.vel.linear.x = r.f64(); o.vel.linear.y = r.f64(); o.vel.linear.z = r.f64();
  o.vel.angular.x = r.f64(); o.vel.angular.y = r.f64(); o.vel.angular.z = r.f64();
}
template <typename T> bool finish(Reader& r, size_t* used) { if (used) *used = r.off; return r.ok; }
}  // namespace

Bytes serialize(const MotionRequest& m) { Writer w; put(w, m); return w.b; }
Bytes serialize(const Trajectory& m) { Writer w; put(w, m); return w.b; }
Bytes serialize(const MyReference& m) { Writer w; put(w, m); return w.b; }
Bytes serialize(const MotionResponse& m) {
  Writer w;
  w.u32((uint32_t)m.ref.size());
  for (const auto& x : m.ref) put(w, x);
  w.u32((uint32_t)m.tra.size());
  for (const auto& x : m.tra) put(w, x);
  return w.b;
}
Bytes serialize(const std::vector<Obstacle2D>& obstacles) {
  Writer w;
  w.u32((uint32_t)obstacles.size());
  for (const auto& o : obstacles) put(w, o);
  return w.b;
}
Bytes serializeState(const std::vector<double>& state) { Writer w; w.f64s(state); return w.b; }

bool deserialize(const uint8_t* p, size_t n, MotionRequest& m, size_t* used) { Reader r{p, n}; get(r, m); return finish<void>(r, used); }
bool deserialize(const uint8_t* p, size_t n, Trajectory& m, size_t* used) { Reader r{p, n}; get(r, m); return finish<void>(r, used); }
bool deserialize(const uint8_t* p, size_t n, MyReference& m, size_t* used) { Reader r{p, n}; get(r, m); return finish<void>(r, used); }
bool deserialize(const uint8_t* p, size_t n, MotionResponse& m, size_t* used) {
  Reader r{p, n};
  uint32_t k = r.u32();
  if (!r.ok || k > kMaxArray) return false;
  m.ref.assign(k, MyReference());
  for (auto& x : m.ref) { get(r, x); if (!r.ok) return false; }
  k = r.u32();
  if (!r.ok || k > kMaxArray) return false;
  m.tra.assign(k, Trajectory());
  for (auto& x : m.tra) { get(r, x); if (!r.ok) return false; }
  return finish<void>(r, used);
}
bool deserialize(const uint8_t* p, size_t n, std::vector<Obstacle2D>& obstacles, size_t* used) {
  Reader r{p, n};
  const uint32_t k = r.u32();
  if (!r.ok || k > kMaxArray || !r.need(88 * (size_t)k)) return false;
  obstacles.assign(k, Obstacle2D());
  for (auto& o : obstacles) get(r, o);
  return finish<void>(r, used);
}
bool deserializeState(const uint8_t* p, size_t n, std::vector<double>& state, size_t* used) {
  Reader r{p, n};
  r.f64s(state);
  return finish<void>(r, used);
}
}  // namespace wire

wire::MotionResponse preparePathMessage(const std::vector<Path>& path) {
  wire::MotionResponse resp;
  for (auto it = path.begin(); it != path.end(); ++it) {
    MyReference ref;
    ref.dir = it->ref.dir;
    ref.x.insert(ref.x.begin(), it->ref.x.begin(), it->ref.x.end());
    ref.y.insert(ref.y.begin(), it->ref.y.begin(), it->ref.y.end());
    ref.v.insert(ref.v.begin(), it->ref.v.begin(), it->ref.v.end());
    resp.ref.push_back(ref);
    Trajectory tra;
    for (size_t i = 1; i < it->tra.size(); i++) {
      tra.x.push_back(it->tra[i][0]); tra.y.push_back(it->tra[i][1]); tra.theta.push_back(it->tra[i][2]);
      tra.delta.push_back(it->tra[i][3]); tra.v.push_back(it->tra[i][4]); tra.a.push_back(it->tra[i][5]);
      tra.a_cmd.push_back(it->tra[i][8]); tra.d_cmd.push_back(it->tra[i][9]);
    }
    resp.tra.push_back(tra);
  }
  return resp;
}

std::vector<Path> getCommittedPath(std::vector<Node> bestPath, double& Tp, double sim_dt, double Tcommit) {
  std::vector<Path> commit;
  for (auto it = bestPath.begin(); it != bestPath.end(); ++it) {
    Path path;
    path.ref.dir = it->ref.dir;
    for (size_t j = 0; j != it->tra.size(); ++j) {
      Tp += sim_dt;
      path.tra.push_back(it->tra[j]);
      int IDwp = (int)it->tra[j][7];
      if (IDwp < 0 || IDwp >= (int)it->ref.x.size()) throw Error("getCommittedPath: waypoint index outside the node's reference");
      path.ref.x.push_back(it->ref.x[IDwp]);
      path.ref.y.push_back(it->ref.y[IDwp]);
      path.ref.v.push_back(it->ref.v[IDwp]);
      if ((Tp >= Tcommit) && (path.ref.x.size() >= 3)) {
        commit.push_back(path);
        return commit;
      }
    }
    commit.push_back(path);
  }
  return commit;
}

}  // namespace clrrt

// ---- flat entry points ---------------------------------------------------------------------------------------------
extern "C" int clrrt_wire_parse_request(const uint8_t* buf, int n, double* goal4, double* vmax, int* bend, int* n_lane_shifts) {
  clrrt::MotionRequest m;
  if (!buf || !goal4 || n < 0 || !clrrt::wire::deserialize(buf, (size_t)n, m) || m.goal.size() < 4) return CLRRT_ERR_ARG;
  for (int i = 0; i < 4; i++) goal4[i] = m.goal[i];
  if (vmax) *vmax = m.vmax;
  if (bend) *bend = m.bend ? 1 : 0;
  if (n_lane_shifts) *n_lane_shifts = (int)m.laneShifts.size();
  return CLRRT_OK;
}
extern "C" int clrrt_wire_parse_state(const uint8_t* buf, int n, double* state6) {
  std::vector<double> s;
  if (!buf || !state6 || n < 0 || !clrrt::wire::deserializeState(buf, (size_t)n, s) || s.size() != 6) return CLRRT_ERR_ARG;
  for (int i = 0; i < 6; i++) state6[i] = s[i];
  return CLRRT_OK;
}
extern "C" int clrrt_wire_parse_obstacles(const uint8_t* buf, int n, clrrt_obstacle* out, int cap) {
  std::vector<clrrt::Obstacle2D> o;
  if (!buf || n < 0 || !clrrt::wire::deserialize(buf, (size_t)n, o)) return CLRRT_ERR_ARG;
  for (int i = 0; i < (int)o.size() && i < cap; i++)
    out[i] = {o[i].obb.center.x, o[i].obb.center.y, o[i].obb.center.theta, o[i].obb.size_x, o[i].obb.size_y,
              o[i].vel.linear.x, o[i].vel.linear.y};
  return (int)o.size();
}
extern "C" int clrrt_wire_trajectory(const double* rows8, int n_rows, uint8_t* out, int cap) {
  clrrt::Trajectory t;
  for (int i = 0; i < n_rows; i++) {
    const double* r = rows8 + 8 * i;
    t.x.push_back(r[0]); t.y.push_back(r[1]); t.theta.push_back(r[2]); t.delta.push_back(r[3]); t.v.push_back(r[4]);
    t.a.push_back(r[5]); t.a_cmd.push_back(r[6]); t.d_cmd.push_back(r[7]);
  }
  const clrrt::wire::Bytes b = clrrt::wire::serialize(t);
  if ((int)b.size() > cap) return CLRRT_ERR_CAPACITY;
  memcpy(out, b.data(), b.size());
  return (int)b.size();
}
extern "C" int clrrt_wire_obstacles(const clrrt_obstacle* obs, int n, uint8_t* out, int cap) {
  std::vector<clrrt::Obstacle2D> v((size_t)n);
  for (int i = 0; i < n; i++) {
    v[i].obb.center.x = obs[i].cx; v[i].obb.center.y = obs[i].cy; v[i].obb.center.theta = obs[i].theta;
    v[i].obb.size_x = obs[i].size_x; v[i].obb.size_y = obs[i].size_y; v[i].vel.linear.x = obs[i].vx; v[i].vel.linear.y = obs[i].vy;
  }
  const clrrt::wire::Bytes b = clrrt::wire::serialize(v);
  if ((int)b.size() > cap) return CLRRT_ERR_CAPACITY;
  memcpy(out, b.data(), b.size());
  return (int)b.size();
}
// plan segments (rows of 10 state entries) -> preparePathMessage -> MotionResponse bytes -> parsed back and re-serialised;
// returns the byte count, or an error if the round trip is not the identity
extern "C" int clrrt_wire_response_roundtrip(const double* rows10, const int32_t* rows_per_segment, int n_segments, uint8_t* out, int cap) {
  std::vector<clrrt::Path> plan((size_t)n_segments);
  size_t row = 0;
  for (int s = 0; s < n_segments; s++) {
    for (int i = 0; i < rows_per_segment[s]; i++, row++) {
      plan[s].tra.emplace_back(rows10 + 10 * row, rows10 + 10 * (row + 1));
      plan[s].ref.x.push_back(rows10[10 * row]); plan[s].ref.y.push_back(rows10[10 * row + 1]); plan[s].ref.v.push_back(rows10[10 * row + 4]);
    }
    plan[s].ref.dir = 1;
  }
  const clrrt::wire::MotionResponse resp = clrrt::preparePathMessage(plan);
  const clrrt::wire::Bytes b = clrrt::wire::serialize(resp);
  clrrt::wire::MotionResponse back;
  size_t used = 0;
  if (!clrrt::wire::deserialize(b.data(), b.size(), back, &used) || used != b.size()) return CLRRT_ERR_STATE;
  if (clrrt::wire::serialize(back) != b) return CLRRT_ERR_STATE;
  if ((int)b.size() > cap) return CLRRT_ERR_CAPACITY;
  memcpy(out, b.data(), b.size());
  return (int)b.size();
}
