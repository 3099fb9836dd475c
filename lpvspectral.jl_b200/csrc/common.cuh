// Shared device helpers for liblpvs (sm_100a).  FP64 throughout.
//
// Tensor path: on sm_100a every mma.sync f64 shape lowers to DMMA.8x8x4 (measured 36.96 TFLOP/s on B200, the same
// pipe as DFMA at 36.7 TFLOP/s -- tools/fp64_probe.cu), tcgen05 has no f64 kind.  So the MMA core below is a
// warp-level m8n8k4 kernel fed from shared memory, and everything around it is built to spend as few other FP64
// issue slots as possible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lpvs {

constexpr int TB = 128;      // tile edge (rows/cols of G per CTA)
constexpr int KC = 32;       // samples (k-dimension) per pipeline chunk
constexpr int LDT = KC + 4;  // smem row stride in doubles: == 4 (mod 16) -> conflict-free DMMA fragment loads
constexpr int TILE_D = TB * LDT;
constexpr int NTHREADS = 256;  // 8 warps: 4 (M) x 2 (N), warp tile 32 x 64
constexpr int FB = 64;         // frequencies per tile block (64 cos + 64 sin columns)
constexpr int GRP = 8;         // chain length: one exact anchor every GRP frequencies

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// one [TB x KC] operand tile (row-major source, leading dimension ld) into a [TB][LDT] smem tile, 16 bytes per copy
__device__ __forceinline__ void load_tile_async(double* dst, const double* src, long long ld, int tid) {
#pragma unroll
    for (int i = 0; i < (TB * KC / 2) / NTHREADS; i++) {
        int q = tid + i * NTHREADS;
        int row = q >> 4, seg = q & 15;
        cp_async16(dst + row * LDT + 2 * seg, src + (long long)row * ld + 2 * seg);
    }
}

// ---- mbarrier (shared-memory barrier object): arrive is non-blocking, waiting does not count as arrival ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra WAIT_DONE;\n"
        " bra WAIT_LOOP;\n WAIT_DONE:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}

// One k4-step of the 32x64 warp tile: A fragments from sA (rows = M), B fragments from sB (rows = N),
// both stored [row][k] with stride LDT.
__device__ __forceinline__ void mma_step(const double* __restrict__ pa, const double* __restrict__ pb, int kk,
                                         double (&acc)[4][8][2]) {
    double a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = pa[i * 8 * LDT + 4 * kk];
#pragma unroll
    for (int j = 0; j < 8; j++) b[j] = pb[j * 8 * LDT + 4 * kk];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
}

// tile index -> (I,J) with I >= J, row-major over the lower triangle
__device__ __forceinline__ void tile_ij(int id, int& I, int& J) {
    int i = (int)((sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= id) i++;
    while (i * (i + 1) / 2 > id) i--;
    I = i;
    J = id - i * (i + 1) / 2;
}

// exact fractional part of f*t in turns, then (cos, -sin)(2*pi*f*t)
__device__ __forceinline__ double2 cis_turns_exact(double f, double t) {
    double hi = __dmul_rn(f, t);  // no contraction: hi must be the rounded product
    double lo = fma(f, t, -hi);
    double r = (hi - rint(hi)) + lo;
    double s, c;
    sincospi(2.0 * r, &s, &c);
    return make_double2(c, -s);
}

// (cos, -sin) of an angle phi [rad] that is already rounded the reference's way; accurate for |phi| < 1e15.
__device__ __forceinline__ double2 cis_of_phase(double phi) {
    const double INV2PI_HI = 0.15915494309189535;      // fl(1/(2 pi))
    const double INV2PI_LO = -9.839338337591243e-18;  // 1/(2 pi) - INV2PI_HI
    double q = __dmul_rn(phi, INV2PI_HI);
    double e = fma(phi, INV2PI_HI, -q);
    double lo = fma(phi, INV2PI_LO, e);
    double r = (q - rint(q)) + lo;
    double s, c;
    sincospi(2.0 * r, &s, &c);
    return make_double2(c, -s);
}

// reference rounding of the Fourier phase: fl(fl(2pi*f)*t)  (src/lsfft.jl:34,41)
__device__ __forceinline__ double2 cis_reference(double f, double t) {
    const double TWO_PI = 6.283185307179586;
    double w = __dmul_rn(TWO_PI, f);
    double phi = __dmul_rn(w, t);
    return cis_of_phase(phi);
}

// One step of the angle-addition chain with the contraction PINNED (one rounded product, one FMA per component), so every
// tile and every kernel that synthesises a column produces bit-identical values: the Gram matrix and the right-hand side
// are then those of ONE well-defined matrix A~.  (With compiler-chosen contraction the same column differed in the last
// bit between the diagonal and off-diagonal instantiations; in a null direction of A that inconsistency is amplified by
// 1/shift ~ 1e13 -- 1.6e-5 in the Nyquist coefficient of the reference's 1000 x 1001 KAT, profiles/r02_rankdef.md.)
__device__ __forceinline__ double2 chain_rotate(double2 z, double2 d) {
    return make_double2(__fma_rn(z.x, d.x, -__dmul_rn(z.y, d.y)), __fma_rn(z.x, d.y, __dmul_rn(z.y, d.x)));
}

}  // namespace lpvs
