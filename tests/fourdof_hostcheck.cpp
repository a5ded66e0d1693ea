// TEST INFRASTRUCTURE ONLY — never linked into libpgs.so.
// Compiles the DEVICE math header of the alternative functors (solve_keyframe_pose_graph_b200/csrc/pgs_fourdof.cuh)
// for the host, so that tests/test_fourdof.py can compare the arithmetic the GPU kernel runs with the oracle on a
// machine without a GPU.  The kernels themselves (launch, staging, stores) are covered by the -m gpu test.
#include <cstddef>
#define PGS_HD inline
#include "../solve_keyframe_pose_graph_b200/csrc/pgs_fourdof.cuh"

using namespace pgs::fourdof;

extern "C" void hostcheck_fourdof(int kind, const double* rot, const double* t, int n_edges, const int* c1, const int* c2, const double* obs_rot,
                                  const double* obs_t, const double* weight, const double* sw, double* r, double* J) {
  for (int e = 0; e < n_edges; ++e) {
    const int a = c1[e], b = c2[e];
    if (kind == 0) {
      Dual<12> res[6];
      four_dof_error<false, 12>(rot + 4 * a, t + 3 * a, rot + 4 * b, t + 3 * b, obs_rot + 4 * e, obs_t + 3 * e, weight[e], 0.0, res);
      for (int i = 0; i < 6; ++i) { r[6 * e + i] = res[i].a; for (int j = 0; j < 12; ++j) J[(size_t)72 * e + 12 * i + j] = res[i].v[j]; }
    } else if (kind == 1) {
      Dual<13> res[7];
      four_dof_error<true, 13>(rot + 4 * a, t + 3 * a, rot + 4 * b, t + 3 * b, obs_rot + 4 * e, obs_t + 3 * e, 1.0, sw[e], res);
      for (int i = 0; i < 7; ++i) { r[7 * e + i] = res[i].a; for (int j = 0; j < 13; ++j) J[(size_t)91 * e + 13 * i + j] = res[i].v[j]; }
    } else {
      Dual<8> res[4];
      qin_four_dof(rot[3 * a], t + 3 * a, rot[3 * b], t + 3 * b, obs_t + 3 * e, obs_rot[3 * e], obs_rot[3 * e + 1], obs_rot[3 * e + 2], res);
      for (int i = 0; i < 4; ++i) { r[4 * e + i] = res[i].a; for (int j = 0; j < 8; ++j) J[(size_t)32 * e + 8 * i + j] = res[i].v[j]; }
    }
  }
}
