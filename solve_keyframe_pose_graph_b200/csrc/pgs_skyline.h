// K4 (direct variant): block skyline (row-envelope) Cholesky of the reduced pose system on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace pgs {
struct SkylineFactor;
// Symbolic phase (host): envelope of every node row from the pair list (hi > lo), panel partition, row lists.
SkylineFactor* skyline_create(int N, int n_pairs, const int* pair_hi, const int* pair_lo, cudaStream_t stream, std::string* err);
void skyline_destroy(SkylineFactor* f);
int64_t skyline_nnz(const SkylineFactor* f);
// Numeric phase (device): scatter Ad[N][36] / Ao[P][36] into the envelope, factor A = L L^T, solve A y = b.
// Returns PGS_OK, PGS_ERR_LINEAR_SOLVER (non-positive pivot) or a CUDA error code.
int skyline_factor_solve(SkylineFactor* f, const double* Ad, const double* Ao, const double* b, double* y, std::string* err);
}  // namespace pgs
