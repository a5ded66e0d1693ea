// ws_lab — the persistent update kernel under a clock: runs the skyline factorisation on a synthetic band graph of
// config-3 shape (front ~2700 rows) and prints, for the rest kernel of one panel, the cycle stamps of thread 0 of a few
// CTAs: per tile — table + A_old loads issued | per K chunk: operands there, multiply done, barrier + next request | stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DSKY_WS_CLOCKS=60 -o ws_lab tools/ws_lab.cu -ldl
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
  cudaEvent_t e0, e1; CKL(cudaEventCreate(&e0)); CKL(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CKL(cudaEventRecord(e0, st));
    if (skyline_factor(f, dAd, dAo, db, &err)) { fprintf(stderr, "factor: %s\n", err.c_str()); return 1; }
    CKL(cudaEventRecord(e1, st));
    if (skyline_check(f, &err)) { fprintf(stderr, "check: %s\n", err.c_str()); return 1; }
    float a; CKL(cudaEventElapsedTime(&a, e0, e1));
    printf("N=%d panels=%d nnz=%.3g  factor %.2f ms (%.1f us/panel)\n", N, f->D, (double)f->nnz, a, 1e3 * a / f->D);
  }
  static long long clk[64][48];
  CKL(cudaMemcpyFromSymbol(clk, g_ws_clk, sizeof(clk)));
  printf("rest kernel, panel %d, thread 0 of a CTA [cycles since its entry]; per tile: pre = row table + A_old loads issued; per chunk: wait (operands there) / mma / sync+request; st = stores\n", SKY_WS_CLOCKS);
  for (int c = 0; c < 64; c += 9) {
    const long long* k = clk[c];
    if (!k[0]) continue;
    printf("  cta %2d:", c);
    for (int it = 0; it < 3; ++it) {
      if (!k[11 + 12 * it]) break;
      const long long base = it == 0 ? k[0] : k[11 + 12 * (it - 1)];
      printf("  | tile %d pre %5lld", it, k[1 + 12 * it] - base);
      for (int ch = 0; ch < 3; ++ch)
        printf("  c%d %5lld/%5lld/%5lld", ch, k[2 + 12 * it + 3 * ch] - (ch ? k[4 + 12 * it + 3 * (ch - 1)] : k[1 + 12 * it]), k[3 + 12 * it + 3 * ch] - k[2 + 12 * it + 3 * ch],
               k[4 + 12 * it + 3 * ch] - k[3 + 12 * it + 3 * ch]);
      printf("  st %5lld  (tile total %6lld)", k[11 + 12 * it] - k[10 + 12 * it], k[11 + 12 * it] - base);
    }
    printf("\n");
  }
  return 0;
}
