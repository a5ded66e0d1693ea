// upd_lab — runs the skyline factorisation on a synthetic band graph of config-3 shape (envelope ~2700 scalars) and
// prints the cycle stamps of the update kernels' phases for one panel, plus the time of a whole factorisation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSKY_UPD_CLOCKS=60 -o upd_lab tools/upd_lab.cu -ldl
#include "../solve_keyframe_pose_graph_b200/csrc/pgs_skyline.cu"

#include <random>
using namespace pgs;
#define CKL(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 3200, GAP = argc > 2 ? atoi(argv[2]) : 450;
  std::vector<int> hi, lo;
  for (int i = 1; i < N; ++i) { hi.push_back(i); lo.push_back(i - 1); }
  for (int i = GAP; i < N; i += 2) { hi.push_back(i); lo.push_back(i - GAP); }
  const int P = (int)hi.size();
  cudaStream_t st; CKL(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  std::string err;
  SkylineFactor* f = skyline_create(N, P, hi.data(), lo.data(), st, &err);
  if (!f) { fprintf(stderr, "create failed: %s\n", err.c_str()); return 1; }
  std::mt19937_64 rng(3); std::uniform_real_distribution<double> U(-1, 1);
  std::vector<double> Ad((size_t)N * 36, 0.0), Ao((size_t)P * 36), b((size_t)N * 6, 1.0);
  for (int i = 0; i < N; ++i) for (int k = 0; k < 6; ++k) Ad[36 * (size_t)i + 7 * k] = 40.0;
  for (auto& x : Ao) x = 0.5 * U(rng);
  double *dAd, *dAo, *db, *dy;
  CKL(cudaMalloc((void**)&dAd, Ad.size() * 8)); CKL(cudaMalloc((void**)&dAo, Ao.size() * 8)); CKL(cudaMalloc((void**)&db, b.size() * 8)); CKL(cudaMalloc((void**)&dy, b.size() * 8));
  CKL(cudaMemcpy(dAd, Ad.data(), Ad.size() * 8, cudaMemcpyHostToDevice)); CKL(cudaMemcpy(dAo, Ao.data(), Ao.size() * 8, cudaMemcpyHostToDevice));
  CKL(cudaMemcpy(db, b.data(), b.size() * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1, e2; CKL(cudaEventCreate(&e0)); CKL(cudaEventCreate(&e1)); CKL(cudaEventCreate(&e2));
  for (int rep = 0; rep < 2; ++rep) {
    CKL(cudaEventRecord(e0, st));
    if (skyline_factor(f, dAd, dAo, db, &err)) { fprintf(stderr, "factor: %s\n", err.c_str()); return 1; }
    CKL(cudaEventRecord(e1, st));
    if (skyline_backward(f, dy, &err)) { fprintf(stderr, "backward: %s\n", err.c_str()); return 1; }
    CKL(cudaEventRecord(e2, st));
    if (skyline_check(f, &err)) { fprintf(stderr, "check: %s\n", err.c_str()); return 1; }
    float a, c; CKL(cudaEventElapsedTime(&a, e0, e1)); CKL(cudaEventElapsedTime(&c, e1, e2));
    printf("N=%d panels=%d nnz=%.3g  factor %.2f ms (%.1f us/panel)  backward %.2f ms (%.1f us/panel)\n", N, f->D, (double)f->nnz, a, 1e3 * a / f->D, c, 1e3 * c / f->D);
  }
#ifdef SKY_UPD_CLOCKS
  static long long clk[2][64][10];
  CKL(cudaMemcpyFromSymbol(clk, g_upd_clk, sizeof(clk)));
  const char* nm[2] = {"next", "rest"};
  for (int part = 0; part < 2; ++part) {
    long long t0 = clk[part][0][0];
    for (int c = 0; c < 64; ++c) if (clk[part][c][0] && clk[part][c][0] < t0) t0 = clk[part][c][0];
    printf("%s kernel, panel %d: per CTA (group 0): start | index | issue+preload | wait0 mma0 | wait1 mma1 | wait2 mma2 | store  [cycles]\n", nm[part], SKY_UPD_CLOCKS);
    for (int c = 0; c < 64; c += (part ? 7 : 3)) {
      const long long* k = clk[part][c];
      if (!k[0]) continue;
      printf("  cta %2d: +%6lld | %5lld | %5lld | %5lld %5lld | %5lld %5lld | %5lld %5lld | %5lld   total %6lld\n", c, k[0] - t0, k[1] - k[0], k[2] - k[1], k[3] - k[2], k[4] - k[3],
             k[5] - k[4], k[6] - k[5], k[7] - k[6], k[8] - k[7], k[9] - k[8], k[9] - k[0]);
    }
  }
#endif
  return 0;
}
