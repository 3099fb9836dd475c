// Issue-rate probe of the warp-level (legacy) tensor-core MMAs on sm_100a: mma.sync m16n8k8 TF32 and m16n8k16 BF16, FP32
// accumulate, operands in registers, 8 independent accumulator tiles per warp.  What a low-precision correction GEMM written
// with mma.sync (not tcgen05) could reach.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) k_probe(float* out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3u, a2 = threadIdx.x * 5u, a3 = threadIdx.x * 7u, b0 = 11u + threadIdx.x, b1 = 13u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 20000;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 4 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int kind = 0; kind < 2; kind++) {
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (kind == 0) k_probe<0><<<sms * ctas, 256>>>(out, iters);
                else k_probe<1><<<sms * ctas, 256>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double flop = (double)sms * ctas * 8 /*warps*/ * iters * 8.0 * (kind == 0 ? 16.0 * 8 * 8 * 2 : 16.0 * 8 * 16 * 2);
            printf("%s, %d CTA(s) of 8 warps per SM: %.1f TFLOP/s\n", kind == 0 ? "mma.sync m16n8k8 tf32" : "mma.sync m16n8k16 bf16",
                   ctas, flop / (best * 1e-3) / 1e12);
        }
    }
    return 0;
}
