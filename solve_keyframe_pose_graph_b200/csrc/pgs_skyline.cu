#include "pgs_skyline.h"
#include "../../include/pgs.h"
namespace pgs {
struct SkylineFactor { int dummy; };
SkylineFactor* skyline_create(int, int, const int*, const int*, cudaStream_t, std::string* err) { if (err) *err = "skyline Cholesky not built yet"; return nullptr; }
void skyline_destroy(SkylineFactor* f) { delete f; }
int64_t skyline_nnz(const SkylineFactor*) { return 0; }
int skyline_factor_solve(SkylineFactor*, const double*, const double*, const double*, double*, std::string*) { return PGS_ERR_STATE; }
}
