// K4 (direct variant): block skyline (row-envelope) Cholesky of the reduced pose system on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace pgs {
struct SkylineFactor;
// Symbolic phase (host): envelope of every node row from the pair list (hi > lo), panel partition, row lists.
// n_border_nodes > 0: the last n_border_nodes nodes are border unknowns of a domain decomposition — they are
// not eliminated, their rows keep the whole border block, and (N - n_border_nodes) * 6 must be a multiple of
// skyline_panel_width().
// dense: every row starts at column 0 and everything is eliminated (the summed border system of the Schur scheme).
// node_src / pair_src (host, may be null = identity): the factor holds a sub-system of the solver's (Ad, Ao, b) — node i
// reads row node_src[i] of Ad and b (-1 = no diagonal block / rhs from here), pair p reads row pair_src[p] of Ao.
// tail: extra doubles allocated (and zeroed by every numeric phase) right behind the envelope values.
SkylineFactor* skyline_create(int N, int n_pairs, const int* pair_hi, const int* pair_lo, cudaStream_t stream, std::string* err,
                              int n_border_nodes = 0, bool dense = false, const int* node_src = nullptr, const int* pair_src = nullptr,
                              long long tail = 0);
void skyline_destroy(SkylineFactor* f);
int64_t skyline_nnz(const SkylineFactor* f);
void skyline_set_share(SkylineFactor* f, int share);   // how many factorisations run on the GPU at the same time (elimination chains)
int skyline_panel_width();
// Numeric phase (device): scatter Ad[N][36] / Ao[P][36] into the envelope, factor A = L L^T, solve A y = b.
// Returns PGS_OK, PGS_ERR_LINEAR_SOLVER (non-positive pivot) or a CUDA error code.
int skyline_factor_solve(SkylineFactor* f, const double* Ad, const double* Ao, const double* b, double* y, std::string* err);
// The same in pieces (sharded solve): eliminate the interior panels of a chain; add its border Schur complement and
// forward-substituted border rhs into the border factor; after the border solve put x_border into
// y[interior_scalars .. n) and back-substitute the interior; check the pivot flag (synchronises).
int skyline_factor(SkylineFactor* f, const double* Ad, const double* Ao, const double* b, std::string* err);
int skyline_border_accumulate(SkylineFactor* chain, SkylineFactor* border, const int* bmap_dev, cudaStream_t st, std::string* err);
int skyline_backward(SkylineFactor* f, double* y, std::string* err);
int skyline_check(SkylineFactor* f, std::string* err);
int skyline_interior_scalars(const SkylineFactor* f);
int skyline_begin_border(SkylineFactor* f, std::string* err);                  // zero the envelope, the tail and the pivot flag
double* skyline_values(SkylineFactor* f);                                       // envelope values followed by the tail
long long skyline_values_count(const SkylineFactor* f);
double* skyline_tail(SkylineFactor* f);
int skyline_add_diagonal(SkylineFactor* f, const double* add, std::string* err);   // A_rr += add[r]
int skyline_factor_numeric(SkylineFactor* f, std::string* err);                // the panel loop on what is loaded
const int* skyline_fail_flag(const SkylineFactor* f);                                                      // device int, 1 = non-positive pivot
}  // namespace pgs
