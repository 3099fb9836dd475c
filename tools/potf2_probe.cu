// Timing / correctness probe for the 128x128 diagonal-block kernel k_potf2 (chol.cu), outside the library.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DLPVS_POTF2_TRACE \
//        -o tools/potf2_probe tools/potf2_probe.cu
// Run on the GPU box: tools/potf2_probe [nproblems]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../lpvspectral.jl_b200/csrc/chol.cu"

using namespace lpvs;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 2047;
    const int n = 128;
    std::vector<double> A((size_t)n * n), G((size_t)K * n * n);
    srand(1);
    // SPD: B B' + n I
    std::vector<double> B((size_t)n * n);
    for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0;
            for (int k = 0; k < n; k++) s += B[i * n + k] * B[j * n + k];
            A[i * n + j] = s + (i == j ? 4.0 : 0.0);
        }
    for (int p = 0; p < K; p++)
        for (int i = 0; i < n * n; i++) G[(size_t)p * n * n + i] = A[i] * (1.0 + 1e-3 * (p % 7));
    double *dG, *dG0, *dLinv;
    int* dinfo;
    long long* dtrace;
    CK(cudaMalloc(&dG, sizeof(double) * G.size()));
    CK(cudaMalloc(&dG0, sizeof(double) * G.size()));
    CK(cudaMalloc(&dLinv, sizeof(double) * G.size()));
    CK(cudaMalloc(&dinfo, sizeof(int) * K));
    CK(cudaMalloc(&dtrace, sizeof(long long) * 64));
    CK(cudaMemcpy(dG0, G.data(), sizeof(double) * G.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dinfo, 0, sizeof(int) * K));
    CK(cudaMemset(dtrace, 0, sizeof(long long) * 64));
    CholArgs a{};
    a.G = dG; a.strideG = (long long)n * n; a.Linv = dLinv; a.strideLinv = (long long)n * n; a.info = dinfo;
    a.Np = n; a.nb = 1;
#ifdef LPVS_POTF2_TRACE
    CK(cudaMemcpyToSymbol(g_potf2_trace, &dtrace, sizeof(dtrace)));
#endif
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int np : {1, K}) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            CK(cudaMemcpy(dG, dG0, sizeof(double) * G.size(), cudaMemcpyDeviceToDevice));
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            launch_potf2(a, 0, np, 0);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("potf2 x %d: %.1f us  (%.1f us per block-wave)\n", np, best * 1e3, best * 1e3 / ((np + 147) / 148));
    }
    // correctness of problem 0 and K-1
    std::vector<double> L((size_t)n * n), Li((size_t)n * n);
    for (int p : {0, K - 1}) {
        CK(cudaMemcpy(L.data(), dG + (size_t)p * n * n, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(Li.data(), dLinv + (size_t)p * n * n, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
        double e1m = 0, e2m = 0, nrm = 0;
        const double* A0 = G.data() + (size_t)p * n * n;
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) {
                double s = 0, t = 0;
                for (int k = 0; k <= j; k++) s += L[i * n + k] * L[j * n + k];
                for (int k = j; k <= i; k++) t += Li[i * n + k] * L[k * n + j];
                e1m = fmax(e1m, fabs(s - A0[i * n + j]));
                nrm = fmax(nrm, fabs(A0[i * n + j]));
                e2m = fmax(e2m, fabs(t - (i == j ? 1.0 : 0.0)));
            }
        printf("problem %d: |LL'-A|max/|A|max = %.2e, |Linv L - I|max = %.2e\n", p, e1m / nrm, e2m);
    }
    int info0 = 0;
    CK(cudaMemcpy(&info0, dinfo, sizeof(int), cudaMemcpyDeviceToHost));
    printf("info[0] = %d\n", info0);
#ifdef LPVS_POTF2_TRACE
    long long tr[64];
    CK(cudaMemcpy(tr, dtrace, sizeof(tr), cudaMemcpyDeviceToHost));
    printf("block 0 phases [cycles]:");
    for (int i = 1; i < 64 && tr[i]; i++) printf(" %lld", tr[i] - tr[i - 1]);
    printf("\n");
#endif
    return 0;
}
