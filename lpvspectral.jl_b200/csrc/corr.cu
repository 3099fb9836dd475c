// LPVS_PHASE_STRUCTURED_REF: the structured Gram matrix (structured.cu) made parity-grade.
//
// The trigonometric sums give G0 = A0' W A0 for the IDEAL phases theta_ks = 2 pi (f0 + k df) t_s; the reference evaluates its
// basis at phi_ks = fl(fl(2 pi f_k) t_s) (src/lsfft.jl:34,41).  The difference eps_ks = phi_ks - theta_ks (<= 6e-9 rad at
// BASELINE cfg5a) is per-element rounding noise, not a function of i +- j, so it does not fit the Toeplitz + Hankel structure --
// but it is tiny, and to first order (the eps^2 terms are < 1e-17)
//     G_ref = G0 + D'B + B'D,     b_ref = b0 + D'y,     B = A0 (cos, -sin),  D = diag(W) (dA/dphi o eps) = W eps (-sin, -cos).
// dG is 1e-10 of G and is needed to ~3 digits: a HALF-PRECISION tensor-core GEMM (mma.sync m16n8k16, f16 operands, f32
// accumulation) with the operands synthesised in registers -- eps exactly in FP64 (the rounding error of the product by one
// FMA, w t - fl(w t); plus (fl(2 pi f_k) - 2 pi f_k) t), then scaled by a power of two and rounded to f16; the cos / sin factors
// from an FP32 angle-addition chain anchored on the reference phase itself.  Nothing is read from HBM but t, W (and y, u).
//
// k_gram_corr  one CTA per (lower 128 x 128 tile, problem): K' = 2 n (per sample the pair (D_i B_j, B_i D_j)), G += scale * acc
// k_rhs_corr   one CTA per (64-frequency block, problem): b += D'[y u] in FP32
//
// Accuracy (tests/test_gpu_structured.py, tools/corr_emulation.py): dG to 3e-4 of itself; cfg5a window 4166 (phase 2.6e7 rad,
// cond(A) 2e5): 4e-9 -> 2e-11 against the reference-rounded solve.
#include <cuda_fp16.h>

#include "gram.cuh"

namespace lpvs {

namespace {

constexpr int LDH = 72;            // smem row stride in halves: 64 k' + 8 pad = 144 B, conflict-free ldmatrix and STS.32
constexpr int PANEL_H = TB * LDH;  // halves per panel buffer
constexpr int LDM = TB + 1;        // row stride of the FP32 accumulator tile the epilogue stages in shared memory
constexpr double CF = 6755399441055744.0;  // 1.5 * 2^52: x + CF has round(x) in its low word (|x| < 2^31)

// per-problem powers of two: S with |eps| S <= 2^21 for every element, wsc with max|W| wsc < 2^-6 (so |D| < 2^15 in f16), and
// the magic constant / shift that leave the fraction of a phase in turns in the low word of a double
struct CorrScales {
    double S, wsc, unscale;  // unscale = 1 / (S wsc)
    double cq;               // 1.5 * 2^(52 - fb): q + cq has frac(q) 2^fb in its low word
    int shl;                 // 32 - fb
};

__device__ __forceinline__ int exp_above(double x) {  // x < 2^result (x >= 0; zero / denormal -> -1022)
    return ((__double2hiint(x) >> 20) & 0x7ff) - 1022;
}
__device__ __forceinline__ double pow2i(int e) { return __hiloint2double((1023 + e) << 20, 0); }  // |e| <= 1022

__device__ CorrScales corr_scales(const CorrArgs& a, long long s_begin, double* red /* 32 doubles of smem */) {
    const int nwarps = blockDim.x >> 5;
    double tm = 0.0, wm = 0.0;
    for (int idx = threadIdx.x; idx < a.n; idx += blockDim.x) {
        const long long s = s_begin + idx;
        if (s < a.s_end) {
            tm = fmax(tm, fabs(a.t[s]));
            wm = fmax(wm, a.W ? fabs(a.W[a.w_abs ? s : (long long)idx]) : 1.0);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tm = fmax(tm, __shfl_xor_sync(0xffffffffu, tm, o));
        wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = tm;
        red[16 + (threadIdx.x >> 5)] = wm;
    }
    __syncthreads();
    tm = 0.0;
    wm = 0.0;
    for (int i = 0; i < nwarps; i++) {
        tm = fmax(tm, red[i]);
        wm = fmax(wm, red[16 + i]);
    }
    CorrScales sc;
    // |w t - fl(w t)| <= |w t| 2^-53, |dw t| <= dwmax tmax
    const double epsmax = a.wmax * tm * 1.1102230246251565e-16 + a.dwmax * tm;
    const int sexp = min(21 - exp_above(epsmax), 600), wexp = min(-6 - exp_above(wm), 600);
    sc.S = pow2i(sexp);
    sc.wsc = pow2i(wexp);
    sc.unscale = pow2i(-sexp) * pow2i(-wexp);
    const int tb = exp_above(a.wmax * tm * 0.15915494309189535);  // |phase| < 2^tb turns
    const int fb = max(1, min(24, 51 - tb));
    sc.cq = 1.5 * pow2i(52 - fb);
    sc.shl = 32 - fb;
    return sc;
}

// (cos, -sin)(2 pi q) in FP32 for a phase of q turns given as a double (|q| < 2^(51 - fb))
__device__ __forceinline__ float2 cis_turns_f32(double q_plus_cq, int shl) {
    const int fr = (int)((unsigned)__double2loint(q_plus_cq) << shl);  // frac(q) 2^32, signed: [-1/2, 1/2) turns
    const float ang = __int2float_rn(fr) * 1.4629180792671596e-9f;  // 2 pi / 2^32
    return make_float2(__cosf(ang), -__sinf(ang));
}

// The 8 consecutive columns (one chain group) of one sample: emit(j, Bc, Bs, Dc, Ds) with B = (cos, -sin) of the phase and
// D = W eps S wsc (-sin, -cos).  wt: (w, dw S) of the 8 columns.  Columns beyond ncc have w = dw = 0, so their D is zero; their B
// is not (the consumers of G / b skip those rows and columns).  The FP64 part of all 8 elements is issued first (independent
// chains: latency hidden), then the FP32 angle-addition chain.
template <class Emit>
__device__ __forceinline__ void corr_group(const double2* __restrict__ wt, double t, float wg, const CorrScales& sc, float2 step,
                                           Emit&& emit) {
    int ei[GRP];
    const double negS = -sc.S;
    double p0 = 0.0;
#pragma unroll
    for (int j = 0; j < GRP; j++) {
        const double2 wj = wt[j];
        const double p = __dmul_rn(wj.x, t);         // fl(w t)
        const double e = __fma_rn(wj.x, t, -p);      // w t - fl(w t), exact
        const double g = __fma_rn(e, negS, CF);      // fixed point: -(w t - p) S
        ei[j] = __double2loint(__fma_rn(wj.y, t, g));  // + (w - w_ideal) S t  =  eps S
        if (j == 0) p0 = p;
    }
    float2 z = cis_turns_f32(__fma_rn(p0, 0.15915494309189535, sc.cq), sc.shl);  // anchor: the reference's own phase
#pragma unroll
    for (int j = 0; j < GRP; j++) {
        const float ef = __int_as_float(0x4B400000 + ei[j]) - 12582912.0f;  // int -> float, exact below 2^22
        const float de = ef * wg;
        emit(j, z.x, z.y, de * z.y, -de * z.x);
        z = make_float2(fmaf(z.x, step.x, -z.y * step.y), fmaf(z.x, step.y, z.y * step.x));
    }
}

__device__ __forceinline__ void ldsm_x4(unsigned (&r)[4], const __half* p) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(s));
}
__device__ __forceinline__ void hmma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// one sample of a chunk as loaded (the conversion of the weight waits until the chunk is synthesised: the loads run two chunks
// ahead and nothing may depend on them earlier)
struct Sample {
    double t, w;
    bool valid;
};

__device__ __forceinline__ Sample load_sample(const CorrArgs& a, long long s_begin, int c, int lane) {
    const int idx = c * KC + lane;
    Sample r;
    r.valid = idx < a.n && s_begin + idx < a.s_end;
    long long s = s_begin + idx;
    if (!r.valid) s = min(s_begin + (long long)a.n, a.s_end) - 1;
    r.t = __ldg(a.t + s);
    r.w = a.W ? __ldg(a.W + (a.w_abs ? s : (s - s_begin))) : 1.0;
    return r;
}
__device__ __forceinline__ float sample_weight(const Sample& sm, double wsc) { return sm.valid ? (float)(sm.w * wsc) : 0.f; }

constexpr int CORR_THREADS = 512;  // warps 0-7: MMA consumers (4 x 2, warp tile 32 x 64); warps 8-15: operand producers
__device__ __forceinline__ void bar_chunk() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 2, 256;\n" ::: "memory"); }

// Warp-specialised: the producers synthesise chunk c + 1 into one panel buffer while the consumers run the MMAs of chunk c
// from the other; one 512-thread named barrier per chunk publishes a buffer and releases the other.
// Off-diagonal tile (I > J): K' = 2 per sample, P = [D_I | B_I], Q = [B_J | D_J]  ->  acc = D_I'B_J + B_I'D_J.
// Diagonal tile: K' = 1 per sample, P = D_I, Q = B_I -> acc = M = D_I'B_I; the epilogue adds M + M' through shared memory.
__global__ void __launch_bounds__(CORR_THREADS, 1) k_gram_corr(const __grid_constant__ CorrArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sP = reinterpret_cast<__half*>(smem_raw);            // [2][TB][LDH]: rows = functions of block I
    __half* sQ = sP + 2 * PANEL_H;                                // [2][TB][LDH]: rows = functions of block J
    double2* sW = reinterpret_cast<double2*>(sQ + 2 * PANEL_H);  // [2][FB]: (w, dw S) of block I, block J
    double* red = reinterpret_cast<double*>(sW + 2 * FB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int I, J;
    tile_ij(blockIdx.x, I, J);
    const int prob = blockIdx.y;
    const long long s_begin = a.start0 + (long long)prob * a.hop;
    const CorrScales sc = corr_scales(a, s_begin, red);
    if (tid < 2 * FB) {
        const int k = (tid < FB ? I : J) * FB + (tid & (FB - 1));
        double2 v = k < a.ncc ? a.wtab[k] : make_double2(0.0, 0.0);
        v.y *= sc.S;
        sW[tid] = v;
    }
    __syncthreads();
    const bool diag = I == J;
    const int nchunks = (a.n + KC - 1) / KC;
    if (warp >= 8) {
        // ---- producers: warp = chain group (columns 8 g .. 8 g + 7 of a block), lane = sample of the chunk
        const int g8 = GRP * (warp - 8);
        Sample cur = load_sample(a, s_begin, 0, lane), nxt = load_sample(a, s_begin, nchunks > 1 ? 1 : 0, lane);
        for (int c = 0; c < nchunks; c++) {
            __half* P = sP + (c & 1) * PANEL_H + g8 * LDH;
            __half* Q = sQ + (c & 1) * PANEL_H + g8 * LDH;
            const Sample nx2 = load_sample(a, s_begin, c + 2 < nchunks ? c + 2 : c, lane);
            const float wg = sample_weight(cur, sc.wsc);
            const float2 step = cis_turns_f32(__fma_rn(a.df, cur.t, sc.cq), sc.shl);
            if (diag) {
                corr_group(sW + g8, cur.t, wg, sc, step, [&](int j, float bc, float bs, float dc, float ds) {
                    P[j * LDH + lane] = __float2half_rn(dc);
                    P[(j + FB) * LDH + lane] = __float2half_rn(ds);
                    Q[j * LDH + lane] = __float2half_rn(bc);
                    Q[(j + FB) * LDH + lane] = __float2half_rn(bs);
                });
            } else {
                corr_group(sW + g8, cur.t, wg, sc, step, [&](int j, float bc, float bs, float dc, float ds) {
                    *reinterpret_cast<__half2*>(P + j * LDH + 2 * lane) = __floats2half2_rn(dc, bc);
                    *reinterpret_cast<__half2*>(P + (j + FB) * LDH + 2 * lane) = __floats2half2_rn(ds, bs);
                });
                corr_group(sW + FB + g8, cur.t, wg, sc, step, [&](int j, float bc, float bs, float dc, float ds) {
                    *reinterpret_cast<__half2*>(Q + j * LDH + 2 * lane) = __floats2half2_rn(bc, dc);
                    *reinterpret_cast<__half2*>(Q + (j + FB) * LDH + 2 * lane) = __floats2half2_rn(bs, ds);
                });
            }
            bar_chunk();  // chunk c published; the consumers have finished chunk c - 1 (the buffer chunk c + 1 goes to)
            cur = nxt;
            nxt = nx2;
        }
    } else {
        // ---- consumers
        float acc[2][8][4];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;
        // ldmatrix lane addresses: A = P rows (16 x 16: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)), B = Q rows ((n 0-7 | 8-15) x k)
        const int wm = warp & 3, wn = warp >> 2, mi = lane >> 3;
        const int a_off = (wm * 32 + (mi & 1) * 8 + (lane & 7)) * LDH + (mi >> 1) * 8;
        const int b_off = (wn * 64 + (mi >> 1) * 8 + (lane & 7)) * LDH + (mi & 1) * 8;
        const int ksteps = diag ? KC / 16 : 2 * KC / 16;
        for (int c = 0; c < nchunks; c++) {
            bar_chunk();
            const __half* P = sP + (c & 1) * PANEL_H;
            const __half* Q = sQ + (c & 1) * PANEL_H;
            for (int ks = 0; ks < ksteps; ks++) {
                unsigned af[2][4];
#pragma unroll
                for (int i = 0; i < 2; i++) ldsm_x4(af[i], P + a_off + i * 16 * LDH + ks * 16);
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    unsigned bf[4];
                    ldsm_x4(bf, Q + b_off + jj * 16 * LDH + ks * 16);
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        hmma16816(acc[i][2 * jj], af[i], bf[0], bf[1]);
                        hmma16816(acc[i][2 * jj + 1], af[i], bf[2], bf[3]);
                    }
                }
            }
        }
        // the accumulators go through shared memory (over the panel buffers) so that all 16 warps write G, coalesced
        bar_consumers();  // every consumer is done with the panels
        float* sM = reinterpret_cast<float*>(smem_raw);
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int q = 0; q < 4; q++)
                    sM[(wm * 32 + i * 16 + (lane >> 2) + 8 * (q >> 1)) * LDM + wn * 64 + j * 8 + 2 * (lane & 3) + (q & 1)] =
                        acc[i][j][q];
    }
    __syncthreads();
    // G += gscale / (S wsc) * acc (diagonal tiles: M + M'); rows / columns of columns that do not exist (k >= ncc) stay as filled
    {
        const float* sM = reinterpret_cast<const float*>(smem_raw);
        const int Np = a.nblk * TB;
        const double scl = a.gscale * sc.unscale;
        double* Gt = a.G + (long long)prob * a.strideG + (long long)I * TB * Np + J * TB;
        const int cl = tid & (TB - 1);
        const bool cok = J * FB + (cl & (FB - 1)) < a.ncc;
#pragma unroll 8
        for (int r0 = 0; r0 < TB; r0 += CORR_THREADS / TB) {
            const int rl = r0 + (tid >> 7);
            float v = sM[rl * LDM + cl];
            if (diag) v += sM[cl * LDM + rl];
            if (cok && I * FB + (rl & (FB - 1)) < a.ncc) Gt[(long long)rl * Np + cl] += scl * (double)v;
        }
    }
}

// b += D'[y u]: thread = (sample of the chunk, chain group), FP32 accumulation over the chunks, fixed-order lane reduction
__global__ void __launch_bounds__(NTHREADS) k_rhs_corr(const __grid_constant__ CorrArgs a) {
    __shared__ double2 sW[FB];
    __shared__ double red[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int I = blockIdx.x, prob = blockIdx.y;
    const long long s_begin = a.start0 + (long long)prob * a.hop;
    const CorrScales sc = corr_scales(a, s_begin, red);
    if (tid < FB) {
        const int k = I * FB + tid;
        double2 v = k < a.ncc ? a.wtab[k] : make_double2(0.0, 0.0);
        v.y *= sc.S;
        sW[tid] = v;
    }
    __syncthreads();
    const int kv = min(GRP, max(0, a.ncc - (I * FB + GRP * warp)));
    const int nchunks = (a.n + KC - 1) / KC;
    float ac[2][GRP], as[2][GRP];
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int j = 0; j < GRP; j++) ac[r][j] = as[r][j] = 0.f;
    for (int c = 0; c < nchunks; c++) {
        const Sample cur = load_sample(a, s_begin, c, lane);
        const int idx = c * KC + lane;
        long long s = s_begin + idx;
        if (!cur.valid) s = min(s_begin + (long long)a.n, a.s_end) - 1;  // weight 0 there
        const float y0 = (float)a.y[s], y1 = a.nrhs > 1 ? (float)a.u[s] : 0.f;
        const float wg = sample_weight(cur, sc.wsc);
        const float2 step = cis_turns_f32(__fma_rn(a.df, cur.t, sc.cq), sc.shl);
        corr_group(sW + GRP * warp, cur.t, wg, sc, step, [&](int j, float, float, float dc, float ds) {
            ac[0][j] = fmaf(dc, y0, ac[0][j]);
            as[0][j] = fmaf(ds, y0, as[0][j]);
            ac[1][j] = fmaf(dc, y1, ac[1][j]);
            as[1][j] = fmaf(ds, y1, as[1][j]);
        });
    }
    const int Np = a.nblk * TB;
    const double scl = a.bscale * sc.unscale;
    double* Bp = a.B + (long long)prob * a.strideB;
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int j = 0; j < GRP; j++) {
            float vc = ac[r][j], vs = as[r][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vc += __shfl_xor_sync(0xffffffffu, vc, o);
                vs += __shfl_xor_sync(0xffffffffu, vs, o);
            }
            if (lane == 0 && r < a.nrhs && j < kv) {
                double* b = Bp + (long long)r * Np + I * TB + GRP * warp + j;
                b[0] += scl * (double)vc;
                b[FB] += scl * (double)vs;
            }
        }
}

constexpr size_t CORR_SMEM = (size_t)4 * PANEL_H * sizeof(__half) + 2 * FB * sizeof(double2) + 32 * sizeof(double);

}  // namespace

int launch_gram_corr(const CorrArgs& a, int nproblems, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_gram_corr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CORR_SMEM);
        attr = true;
    }
    int launched = 0;
    const int ntiles = a.nblk * (a.nblk + 1) / 2;
    for (int p0 = 0; p0 < nproblems; p0 += 32768) {
        const int np = nproblems - p0 < 32768 ? nproblems - p0 : 32768;
        CorrArgs b = a;
        b.start0 = a.start0 + (long long)p0 * a.hop;
        b.G = a.G + (long long)p0 * a.strideG;
        k_gram_corr<<<dim3(ntiles, np), CORR_THREADS, CORR_SMEM, st>>>(b);
        launched++;
    }
    return launched;
}

int launch_rhs_corr(const CorrArgs& a, int nproblems, cudaStream_t st) {
    int launched = 0;
    for (int p0 = 0; p0 < nproblems; p0 += 32768) {
        const int np = nproblems - p0 < 32768 ? nproblems - p0 : 32768;
        CorrArgs b = a;
        b.start0 = a.start0 + (long long)p0 * a.hop;
        b.B = a.B + (long long)p0 * a.strideB;
        k_rhs_corr<<<dim3(a.nblk, np), NTHREADS, 0, st>>>(b);
        launched++;
    }
    return launched;
}

// (w, dw) per complex column for the ideal grid f0 + k df: w = fl(2 pi f_k) as the reference forms it (src/lsfft.jl:34) and
// dw = w - 2 pi (f0 + k df) in double-double; also max |w| and max |dw| (the bounds the kernels scale eps by)
void structured_ref_wtab(double f0, double df, const double* f, int Nf, int ncol, double* out, double* wmax, double* dwmax) {
    const double P_HI = 6.283185307179586, P_LO = 2.4492935982947064e-16;
    *wmax = 0.0;
    *dwmax = 0.0;
    for (int k = 0; k < ncol; k++) out[2 * k] = out[2 * k + 1] = 0.0;
    for (int k = 0; k < Nf; k++) {
        const double w = P_HI * f[k];
        const double kd = (double)k * df, kd_lo = fma((double)k, df, -kd);
        const double s_hi = f0 + kd, bb = s_hi - f0;
        const double s_lo = ((f0 - (s_hi - bb)) + (kd - bb)) + kd_lo;
        const double ph = P_HI * s_hi, pe = fma(P_HI, s_hi, -ph);
        const double p_lo = fma(P_LO, s_hi, pe) + P_HI * s_lo;
        const double dw = (w - ph) - p_lo;
        out[2 * k] = w;
        out[2 * k + 1] = dw;
        if (fabs(w) > *wmax) *wmax = fabs(w);
        if (fabs(dw) > *dwmax) *dwmax = fabs(dw);
    }
}

}  // namespace lpvs
