/* Plain-C caller of liblpvs.so (include/lpvs.h): the same calls the Julia shim makes with `ccall`.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -Llpvspectral.jl_b200 -llpvs -lm \
 *       -Wl,-rpath,$PWD/lpvspectral.jl_b200 -o /tmp/lpvs_example && /tmp/lpvs_example
 *
 * On a machine without a B200 it prints the window bookkeeping, then reports that lpvs_init failed (there is no CPU
 * fallback) and exits 3.  With a GPU it runs ls_spectral (src/lsfft.jl:62-67) and ls_windowpsd (src/lsfft.jl:112-126)
 * on a sine sampled at sorted random instants and prints the peak bins. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "lpvs.h"

static int cmp(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

int main(void) {
    enum { N = 4096, NF = 256, NW = 8, NFW = 32 };
    const double two_pi = 6.283185307179586;
    printf("liblpvs version %d, %d CUDA device(s)\n", lpvs_version(), lpvs_device_count());
    /* Windows2(y,t,n,noverlap): K = (N-n) div (n-noverlap) + 1, noverlap < 0 means n>>1 (src/windows.jl:27-36) */
    const int n = N / NW;
    const long long K = (long long)lpvs_window_count(N, n, -1);
    printf("N=%d, nw=%d -> n=%d samples per window, %lld windows at 50%% overlap\n", N, NW, n, K);
    if (K != 2 * NW - 1) return 2;

    lpvs_ctx* ctx = NULL;
    int rc = lpvs_init(0, &ctx);
    if (rc != LPVS_OK) {
        printf("lpvs_init failed (%d): %s -- no usable B200, and there is no CPU fallback\n", rc, lpvs_last_error(ctx));
        return 3;
    }
    double *t = malloc(sizeof(double) * N), *y = malloc(sizeof(double) * N), *W = malloc(sizeof(double) * n);
    double f[NF], x[2 * NF], fw[NFW], S[NFW];
    srand(1);
    for (int i = 0; i < N; i++) t[i] = 10.0 * rand() / (double)RAND_MAX;
    qsort(t, N, sizeof(double), cmp);
    for (int i = 0; i < N; i++) y[i] = sin(two_pi * 20.0 * t[i]);
    for (int k = 0; k < NF; k++) f[k] = 0.2 * k; /* zero frequency first (check_freq, src/lsfft.jl:20-24) */
    /* per-window grid: spacing 2 / (window duration ~ 1.25 s), so the window's Gram stays well conditioned */
    for (int k = 0; k < NFW; k++) fw[k] = 1.6 * k;
    for (int i = 0; i < n; i++) W[i] = 0.5 * (1.0 - cos(two_pi * i / (double)(n - 1))); /* DSP.hanning(n) */

    int info = 0;
    rc = lpvs_ls_spectral(ctx, y, t, N, f, NF, NULL, 1e-10, x, &info);
    if (rc) { printf("lpvs_ls_spectral: %d %s\n", rc, lpvs_last_error(ctx)); return 1; }
    int peak = 0;
    for (int k = 1; k < NF; k++)
        if (hypot(x[2 * k], x[2 * k + 1]) > hypot(x[2 * peak], x[2 * peak + 1])) peak = k;
    printf("ls_spectral: peak at f = %.2f Hz (expected 20.00), |x|^2 = %.3f (2 Nf = %d)\n", f[peak],
           x[2 * peak] * x[2 * peak] + x[2 * peak + 1] * x[2 * peak + 1], 2 * NF);

    int64_t Kout = 0;
    rc = lpvs_ls_window(ctx, LPVS_WIN_PSD, y, NULL, t, N, fw, NFW, W, n, -1, 1e-10, S, &Kout, &info);
    if (rc) { printf("lpvs_ls_window: %d %s\n", rc, lpvs_last_error(ctx)); return 1; }
    peak = 0;
    for (int k = 1; k < NFW; k++)
        if (S[k] > S[peak]) peak = k;
    printf("ls_windowpsd (hanning, %lld windows): peak at f = %.1f Hz (nearest bins to 20 Hz: 19.2 / 20.8); %lld kernels "
           "launched so far\n", (long long)Kout, fw[peak], (long long)lpvs_launch_count(ctx));
    lpvs_destroy(ctx);
    free(t); free(y); free(W);
    return 0;
}
