// Deterministic synthetic Manhattan-style 6-DOF pose graphs (include/pgs_synth.h).
#include "../../../include/pgs_synth.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "pose_math.h"

namespace {

struct Rng {  // splitmix64
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
  double uniform(double a, double b) { return a + (b - a) * uniform(); }
  int64_t range(int64_t lo, int64_t hi) { return lo + (int64_t)(uniform() * (double)(hi - lo + 1)); }   // inclusive
  double normal() { double u1 = uniform(), u2 = uniform(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2); }
};

using pgs::Matrix4d;

// Exp of a rotation vector (full angle) into a 4x4
Matrix4d exp_rot(double rx, double ry, double rz) {
  const double th = std::sqrt(rx * rx + ry * ry + rz * rz);
  double q[4];
  if (th < 1e-12) { q[0] = 0.5 * rx; q[1] = 0.5 * ry; q[2] = 0.5 * rz; q[3] = 1.0; }
  else { const double s = std::sin(0.5 * th) / th; q[0] = s * rx; q[1] = s * ry; q[2] = s * rz; q[3] = std::cos(0.5 * th); }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (double& v : q) v /= n;
  Matrix4d T = Matrix4d::Identity();
  pgs::quat_to_rot(q, T);
  return T;
}
Matrix4d noise_pose(Rng& r, double st, double sr) {
  Matrix4d T = exp_rot(sr * r.normal(), sr * r.normal(), sr * r.normal());
  T(0, 3) = st * r.normal(); T(1, 3) = st * r.normal(); T(2, 3) = st * r.normal();
  return T;
}
// re-orthonormalise through the quaternion so long products stay rigid
Matrix4d renorm(const Matrix4d& T) {
  double q[4], t[3];
  pgs::mat_to_raw_xyzw(T, q, t);
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (double& v : q) v /= n;
  return pgs::raw_xyzw_to_mat(q, t);
}

}  // namespace

struct pgs_synth_s {
  std::vector<int64_t> stamps;
  std::vector<double> q, t, gq, gt;
  std::vector<int32_t> a, b; std::vector<double> lq, lt, lw; std::vector<uint8_t> lout;
  std::vector<int64_t> k0, k1;
};

extern "C" {

int pgs_synth_config(int32_t config, pgs_synth_spec* s) {
  if (!s || config < 1 || config > 5) return -1;
  std::memset(s, 0, sizeof(*s));
  s->n_worlds = 1; s->loop_gap_min = 50; s->loop_gap_max = 0;
  s->odom_sigma_t = 0.02; s->odom_sigma_r = 0.002; s->loop_sigma_t = 0.01; s->loop_sigma_r = 0.001;
  s->seed = 0xC0FFEE00ull + (uint64_t)config;
  switch (config) {
    case 1: s->n_nodes = 50; s->n_loop = 1; break;
    case 2: s->n_nodes = 10000; s->n_loop = 2000; break;
    case 3: s->n_nodes = 100000; s->n_loop = 50000; s->outlier_fraction = 0.10; break;
    case 4: s->n_nodes = 25000; s->n_worlds = 4; s->n_loop = 0; s->n_interworld = 200; s->deadzone_nodes = 5; break;
    case 5: s->n_nodes = 1000000; s->n_loop = 500000; s->outlier_fraction = 0.10; break;
  }
  return 0;
}

int pgs_synth_create(const pgs_synth_spec* sp, pgs_synth_handle* out) {
  if (!sp || !out || sp->n_nodes < 2 || sp->n_worlds < 1) return -1;
  pgs_synth_s* G = new pgs_synth_s();
  Rng rng(sp->seed);
  const int W = sp->n_worlds, Nw = sp->n_nodes, DZ = W > 1 ? sp->deadzone_nodes : 0;
  const int N = W * Nw + (W - 1) * DZ;
  G->stamps.resize(N); G->q.resize(4 * (size_t)N); G->t.resize(3 * (size_t)N); G->gq.resize(4 * (size_t)N); G->gt.resize(3 * (size_t)N);
  const int64_t dt_ns = 100000000;  // 10 Hz keyframes
  std::vector<int> world_of(N, -1), world_start(W), world_end(W);
  // ---- ground truth: 1 m per node along body +x; every 200 nodes a +-90 deg yaw turn spread over 180 nodes
  // (0.5 deg/node); every 1000 nodes a +-2 m z-ramp over 100 nodes (sinusoidal pitch, peak 1.8 deg);
  // roll wiggle 0.2 deg * sin(i/50).
  std::vector<Matrix4d> GT(N);
  {
    Rng tr(sp->seed ^ 0xA5A5A5A5ull);
    double yaw = 0.0; int ysign = 1, psign = 1;
    double px = 0, py = 0, pz = 0;
    for (int i = 0; i < N; ++i) {
      if (i % 200 == 0) ysign = (tr.next() & 1) ? 1 : -1;
      if (i % 1000 == 0) psign = (tr.next() & 1) ? 1 : -1;
      const int j = i % 1000;
      const double pitch_deg = (j < 100) ? psign * 1.8 * std::sin(M_PI * j / 100.0) : 0.0;
      const double roll_deg = 0.2 * std::sin(i / 50.0);
      const double ypr[3] = {yaw, pitch_deg, roll_deg};
      Matrix4d T = pgs::ypr2R(ypr);
      T(0, 3) = px; T(1, 3) = py; T(2, 3) = pz;
      GT[i] = T;
      px += T(0, 0); py += T(1, 0); pz += T(2, 0);     // advance 1 m along body x
      if (i % 200 < 180) yaw += ysign * 0.5;
    }
  }
  // ---- worlds, dead zones, timestamps, kidnap intervals
  {
    int i = 0;
    for (int w = 0; w < W; ++w) {
      world_start[w] = i;
      for (int k = 0; k < Nw; ++k) world_of[i++] = w;
      world_end[w] = i - 1;
      if (w + 1 < W) for (int k = 0; k < DZ; ++k) world_of[i++] = -(w + 1);
    }
    for (int k = 0; k < N; ++k) G->stamps[k] = 1000000000ll + dt_ns * k;
    for (int w = 0; w + 1 < W; ++w) {   // kidnapped just after the last node of world w, un-kidnapped just before world w+1
      G->k0.push_back(G->stamps[world_end[w]]);   // exactly the last keyframe of the world, so nodeidx_of_world_i_ended() finds it (1 ms rule)
      G->k1.push_back(G->stamps[world_start[w + 1]] - dt_ns / 4);
    }
  }
  // ---- odometry = GT relative motion (x) noise, integrated; each world restarts in its own frame
  {
    Matrix4d M = GT[0];
    for (int i = 0; i < N; ++i) {
      const int w = world_of[i];
      if (i > 0) {
        if (w >= 1 && i == world_start[w]) {
          // new world: VIO restarts in an arbitrary frame (offset up to 100 m, random yaw)
          const double ypr[3] = {rng.uniform(-180.0, 180.0), 0.0, 0.0};
          M = pgs::ypr2R(ypr);
          M(0, 3) = rng.uniform(-100.0, 100.0); M(1, 3) = rng.uniform(-100.0, 100.0); M(2, 3) = rng.uniform(-10.0, 10.0);
        } else {
          Matrix4d rel = GT[i - 1].inverse() * GT[i];
          M = renorm(M * rel * noise_pose(rng, sp->odom_sigma_t, sp->odom_sigma_r));
        }
      }
      pgs::mat_to_raw_xyzw(M, &G->q[4 * (size_t)i], &G->t[3 * (size_t)i]);
      pgs::mat_to_raw_xyzw(GT[i], &G->gq[4 * (size_t)i], &G->gt[3 * (size_t)i]);
    }
  }
  // ---- loop edges (a, b): b ~ U, a = b + gap, observation b_T_a = GT (x) noise or gross outlier
  auto push_edge = [&](int a, int b, bool outlier) {
    Matrix4d bTa = GT[b].inverse() * GT[a];
    if (outlier) {
      double ax = rng.normal(), ay = rng.normal(), az = rng.normal(); double n = std::sqrt(ax * ax + ay * ay + az * az) + 1e-300;
      const double ang = rng.uniform(0.5, M_PI);
      Matrix4d E = exp_rot(ang * ax / n, ang * ay / n, ang * az / n);
      double dx = rng.normal(), dy = rng.normal(), dz = rng.normal(); n = std::sqrt(dx * dx + dy * dy + dz * dz) + 1e-300;
      const double len = rng.uniform(5.0, 50.0);
      E(0, 3) = len * dx / n; E(1, 3) = len * dy / n; E(2, 3) = len * dz / n;
      bTa = bTa * E;
    } else {
      bTa = bTa * noise_pose(rng, sp->loop_sigma_t, sp->loop_sigma_r);
    }
    bTa = renorm(bTa);
    double q[4], t[3];
    pgs::mat_to_raw_xyzw(bTa, q, t);
    G->a.push_back(a); G->b.push_back(b);
    G->lq.insert(G->lq.end(), q, q + 4); G->lt.insert(G->lt.end(), t, t + 3);
    G->lw.push_back(1.0); G->lout.push_back(outlier ? 1 : 0);
  };
  if (Nw == 50 && sp->n_loop == 1 && W == 1) {
    push_edge(49, 0, false);    // config 1: the plumbing case, loop (49 -> 0)
  } else {
    int gmax = sp->loop_gap_max > 0 ? sp->loop_gap_max : std::min(Nw / 8, 2000);
    int gmin = std::min(sp->loop_gap_min, std::max(1, gmax));
    gmax = std::max(gmax, gmin);
    for (int e = 0; e < sp->n_loop; ++e) {
      const int w = (int)rng.range(0, W - 1);
      const int gap = (int)rng.range(gmin, gmax);
      const int b = world_start[w] + (int)rng.range(0, std::max(0, Nw - 1 - gap));
      const int a = std::min(b + gap, world_end[w]);
      push_edge(a, b, rng.uniform() < sp->outlier_fraction);
    }
  }
  // inter-world edges: first one per adjacent world pair, then random pairs of distinct worlds
  for (int e = 0; e < sp->n_interworld && W > 1; ++e) {
    int wa, wb;
    if (e < W - 1) { wb = e; wa = e + 1; }
    else { wb = (int)rng.range(0, W - 1); wa = (int)rng.range(0, W - 2); if (wa >= wb) ++wa; }
    const int b = world_start[wb] + (int)rng.range(0, Nw - 1), a = world_start[wa] + (int)rng.range(0, Nw - 1);
    push_edge(a, b, false);
  }
  *out = G;
  return 0;
}

void pgs_synth_destroy(pgs_synth_handle h) { delete h; }

void pgs_synth_sizes(pgs_synth_handle h, int32_t* n_nodes, int32_t* n_loop, int32_t* n_kidnaps) {
  if (n_nodes) *n_nodes = (int32_t)h->stamps.size();
  if (n_loop) *n_loop = (int32_t)h->a.size();
  if (n_kidnaps) *n_kidnaps = (int32_t)h->k0.size();
}

void pgs_synth_copy(pgs_synth_handle h, int64_t* stamps_ns, double* q, double* t, double* gt_q, double* gt_t, int32_t* a, int32_t* b,
                    double* q_bTa, double* t_bTa, double* weight, uint8_t* is_outlier, int64_t* kidnap_start_ns, int64_t* kidnap_end_ns) {
  auto cp = [](void* dst, const void* src, size_t bytes) { if (dst && bytes) std::memcpy(dst, src, bytes); };
  cp(stamps_ns, h->stamps.data(), 8 * h->stamps.size()); cp(q, h->q.data(), 8 * h->q.size()); cp(t, h->t.data(), 8 * h->t.size());
  cp(gt_q, h->gq.data(), 8 * h->gq.size()); cp(gt_t, h->gt.data(), 8 * h->gt.size());
  cp(a, h->a.data(), 4 * h->a.size()); cp(b, h->b.data(), 4 * h->b.size()); cp(q_bTa, h->lq.data(), 8 * h->lq.size()); cp(t_bTa, h->lt.data(), 8 * h->lt.size());
  cp(weight, h->lw.data(), 8 * h->lw.size()); cp(is_outlier, h->lout.data(), h->lout.size());
  cp(kidnap_start_ns, h->k0.data(), 8 * h->k0.size()); cp(kidnap_end_ns, h->k1.data(), 8 * h->k1.size());
}

}  // extern "C"
