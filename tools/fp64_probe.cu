// FP64 pipe probe for sm_100a: how fast are DMMA (mma.sync m8n8k4 f64) and DFMA, and do they overlap?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_probe tools/fp64_probe.cu
// Run on the GPU box; prints one JSON object. Numbers feed DESIGN.md (FP64 roofline denominator).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NM DMMAs and NF DFMAs interleaved per inner step.
template <int NM, int NF>
__global__ void k_mixed(double* out, int iters, double a0, double b0) {
    double c[NM][2];
    double d[NF];
#pragma unroll
    for (int i = 0; i < NM; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NF; i++) d[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) {
            dmma884(c[i][0], c[i][1], a, b);
            if (i < NF) d[i] = fma(d[i], a, b);
        }
#pragma unroll
        for (int i = NM; i < NF; i++) d[i] = fma(d[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP32 FFMA alongside DMMA (does integer/fp32 work run free under DMMA?)
template <int NM, int NF>
__global__ void k_mixed32(double* out, int iters, double a0, double b0) {
    double c[NM][2];
    float d[NF];
#pragma unroll
    for (int i = 0; i < NM; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NF; i++) d[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    float af = (float)a, bf = (float)b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) {
            dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < NF / NM; j++) d[i * (NF / NM) + j] = fmaf(d[i * (NF / NM) + j], af, bf);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; i++) s += d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The real warp tile: 32 accumulators (4 A-fragments x 8 B-fragments), operands from registers ...
__global__ void k_dmma_tile_reg(double* out, int iters, double a0, double b0) {
    double c[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) c[i][j][0] = c[i][j][1] = 0.0;
    double a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = a0 + i + threadIdx.x * 1e-9;
#pragma unroll
    for (int j = 0; j < 8; j++) b[j] = b0 + j * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ... and from shared memory with the kernel's conflict-free [row][k] stride-36 layout (12 LDS.64 per 32 DMMA)
__global__ void k_dmma_tile_lds(double* out, int iters, double a0, double b0) {
    __shared__ double sA[128 * 36];  // one tile serves as both operands (48 KB static limit)
    const double* sB = sA;
    for (int i = threadIdx.x; i < 128 * 36; i += blockDim.x) sA[i] = a0 + b0 * i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wm = w & 3, wn = (w >> 2) & 1;
    const double* pa = sA + (32 * wm + (lane >> 2)) * 36 + (lane & 3);
    const double* pb = sB + (64 * wn + (lane >> 2)) * 36 + (lane & 3);
    double c[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) c[i][j][0] = c[i][j][1] = 0.0;
    for (int it = 0; it < iters; it += 8) {
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = pa[i * 8 * 36 + 4 * kk];
#pragma unroll
            for (int j = 0; j < 8; j++) b[j] = pb[j * 8 * 36 + 4 * kk];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 8; j++) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int iters = 4096;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    // warps per SM sweep for DMMA
    for (int wps : {4, 8, 16, 32}) {
        int threads = 32 * wps; if (threads > 1024) threads = 1024;
        int blocks = sms * ((32 * wps + threads - 1) / threads);
        float ms = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        double flop = (double)blocks * (threads / 32) * iters * 8 * 512.0;
        printf(", \"dmma_tflops_w%d\": %.3f", wps, flop / ms * 1e-9);
    }
    for (int wps : {4, 8, 16, 32}) {
        int threads = 32 * wps; if (threads > 1024) threads = 1024;
        int blocks = sms * ((32 * wps + threads - 1) / threads);
        float ms = time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        double flop = (double)blocks * threads * iters * 16 * 2.0;
        printf(", \"dfma_tflops_w%d\": %.3f", wps, flop / ms * 1e-9);
    }
    {   // mixed: 8 DMMA + 8 DFMA per step (DFMA work = 8*64 flop vs DMMA 8*512 flop per warp)
        int threads = 512, blocks = sms * 2;
        float ms_m = time_ms([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        float ms_x8 = time_ms([&] { k_mixed<8, 8><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        float ms_x16 = time_ms([&] { k_mixed<8, 16><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        float ms_x32 = time_ms([&] { k_mixed<8, 32><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        float ms_f8 = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        float ms_f32 = time_ms([&] { k_dfma<32><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        float ms_s32 = time_ms([&] { k_mixed32<8, 32><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        float ms_s64 = time_ms([&] { k_mixed32<8, 64><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
        printf(", \"mix_ms\": {\"dmma8\": %.4f, \"dmma8+dfma8\": %.4f, \"dmma8+dfma16\": %.4f, \"dmma8+dfma32\": %.4f, \"dfma8\": %.4f, \"dfma32\": %.4f, \"dmma8+ffma32\": %.4f, \"dmma8+ffma64\": %.4f}",
               ms_m, ms_x8, ms_x16, ms_x32, ms_f8, ms_f32, ms_s32, ms_s64);
    }
    {   // the real 32-accumulator warp tile, 8 warps per SM (256 threads, 1 CTA/SM)
        int threads = 256, blocks = sms;
        int it2 = 8192;
        float ms_r = time_ms([&] { k_dmma_tile_reg<<<blocks, threads>>>(out, it2, 1.0000001, 0.999999); }, 5);
        float ms_l = time_ms([&] { k_dmma_tile_lds<<<blocks, threads>>>(out, it2, 1.0000001, 0.999999); }, 5);
        double flop = (double)blocks * (threads / 32) * it2 * 32 * 512.0;
        printf(", \"tile32_reg_tflops\": %.3f, \"tile32_lds_tflops\": %.3f", flop / ms_r * 1e-9, flop / ms_l * 1e-9);
        threads = 512; blocks = sms;
        ms_r = time_ms([&] { k_dmma_tile_reg<<<blocks, threads>>>(out, it2, 1.0000001, 0.999999); }, 5);
        ms_l = time_ms([&] { k_dmma_tile_lds<<<blocks, threads>>>(out, it2, 1.0000001, 0.999999); }, 5);
        flop = (double)blocks * (threads / 32) * it2 * 32 * 512.0;
        printf(", \"tile32_reg_tflops_16w\": %.3f, \"tile32_lds_tflops_16w\": %.3f", flop / ms_r * 1e-9, flop / ms_l * 1e-9);
    }
    printf("}\n");
    return 0;
}
