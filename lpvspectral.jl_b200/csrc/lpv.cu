// placeholder until the LPV path lands
#include "ctx.h"
extern "C" {
int lpvs_ls_spectral_lpv(lpvs_ctx* c, const double*, const double*, const double*, int64_t, const double*, int, int, double, int, int, double*, double*, double*, int*) { return lpvs::fail(c, LPVS_E_UNSUPPORTED, "lpv not built yet"); }
int64_t lpvs_packed_size(int Nf) { int nb = (Nf + 63) / 64; long long Np = 128LL * nb; return Np * Np + 2 * Np; }
int lpvs_gram_partial_dev(lpvs_ctx* c, const double*, const double*, const double*, const double*, int64_t, const double*, int, double*) { return lpvs::fail(c, LPVS_E_UNSUPPORTED, "not built yet"); }
int lpvs_solve_packed_dev(lpvs_ctx* c, double*, const double*, int, int, double, double*, int*) { return lpvs::fail(c, LPVS_E_UNSUPPORTED, "not built yet"); }
}
