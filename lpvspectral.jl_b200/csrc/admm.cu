// placeholder until the ADMM loop lands (next commit)
#include "ctx.h"
extern "C" {
int lpvs_admm_create_fourier(lpvs_ctx* c, const double*, const double*, int64_t, const double*, int, const double*, int, double, double, const double*, int, double, lpvs_admm**) { return lpvs::fail(c, LPVS_E_UNSUPPORTED, "admm not built yet"); }
int lpvs_admm_create_lpv(lpvs_ctx* c, const double*, const double*, const double*, int64_t, const double*, int, int, int, int, double, double, lpvs_admm**) { return lpvs::fail(c, LPVS_E_UNSUPPORTED, "admm not built yet"); }
int lpvs_admm_run(lpvs_admm*, int64_t, double, int64_t*, double*, int*) { return LPVS_E_UNSUPPORTED; }
int lpvs_admm_size(const lpvs_admm*) { return 0; }
int lpvs_admm_get(lpvs_admm*, double*, double*) { return LPVS_E_UNSUPPORTED; }
int lpvs_admm_result(lpvs_admm*, double*) { return LPVS_E_UNSUPPORTED; }
int lpvs_admm_last_timing(const lpvs_admm*, double*, double*) { return LPVS_E_UNSUPPORTED; }
void lpvs_admm_free(lpvs_admm*) {}
int lpvs_ls_sparse_spectral(lpvs_ctx* c, const double*, const double*, int64_t, const double*, int, const double*, int, double, double, int, double, int64_t, double, double*, int64_t*, double*) { return lpvs::fail(c, LPVS_E_UNSUPPORTED, "admm not built yet"); }
int lpvs_ls_sparse_spectral_lpv(lpvs_ctx* c, const double*, const double*, const double*, int64_t, const double*, int, int, int, int, double, double, int64_t, double, double*, int64_t*, double*) { return lpvs::fail(c, LPVS_E_UNSUPPORTED, "admm not built yet"); }
}
