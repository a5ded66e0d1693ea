// TEST INFRASTRUCTURE ONLY: the product's double printer (csrc/host/Grisu2.h) behind a C function, so that
// tests/test_reference_json.py can compare it value by value with the reference's JSON library.
#include "../solve_keyframe_pose_graph_b200/csrc/host/Grisu2.h"
extern "C" int ours_dump_doubles(int n, const double* v, char* out, int stride) {   // out: n fixed-width records
  for (int i = 0; i < n; ++i) { const std::string s = pgs::grisu2::to_string(v[i]); if ((int)s.size() + 1 > stride) return -1; std::memcpy(out + (size_t)i * stride, s.c_str(), s.size() + 1); }
  return 0;
}
