// LPVS_PHASE_STRUCTURED: the Gram matrix of a Fourier basis on a UNIFORM frequency grid from its trigonometric sums.
//
// With a_k = 2 pi (f0 + k df) t the columns of the regressor are cos a_k and -sin a_k (src/lsfft.jl:26-49), and every entry
// of A' diag(W) A is a product-to-sum identity away from
//     Z-(m) = sum_s W_s e^{-i 2 pi (m df) t_s},          m = 0 .. Nf-1      (differences i - j)
//     Z+(m) = sum_s W_s e^{-i 2 pi (2 f0 + m df) t_s},   m = 0 .. 2Nf-2     (sums i + j):
//     cos_i . cos_j   =  ( Re Z-(i-j) + Re Z+(i+j) ) / 2          (-sin_i) . (-sin_j) = ( Re Z-(i-j) - Re Z+(i+j) ) / 2
//     (-sin_i) . cos_j =  ( Im Z+(i+j) + Im Z-(i-j) ) / 2          cos_i . (-sin_j)    = ( Im Z+(i+j) - Im Z-(i-j) ) / 2
// (Z-(-m) = conj Z-(m)).  The matrix is Toeplitz + Hankel in the frequency index: 3 Nf complex sums of n terms instead of
// Nreg (Nreg + 1) / 2 inner products -- O(n Nf) work per problem where the DMMA kernel of gram.cu does O(n Nf^2).
//
// The phases are those of the IDEAL grid f0 + k df (double-double turns, exact to 1e-32): the class of LPVS_PHASE_CHAIN, not
// the reference's fl(fl(2 pi f) t) -- that rounding is not a function of i +- j, which is why this is an opt-in mode and the
// default stays the DMMA kernel with the reference's phase.
//
// k_sum_tables   per-sample table rows e^{-i 2 pi F_r t_s} for the block anchors of Z-, Z+ and the 8 group powers
// k_trig_sums    one CTA per (64 sums, problem): warp = 8 consecutive m (angle-addition chain), lane = sample of a chunk
// k_gram_fill    one CTA per (lower 128 x 128 tile, problem): the identities above, internal column layout of gram.cuh
#include "gram.cuh"

namespace lpvs {

namespace {

// (cos, -sin)(2 pi F t) for F = hi + lo (double-double), phase reduced exactly in turns
__device__ __forceinline__ double2 cis_turns_exact_dd(double hi, double lo, double t) {
    const double p = __dmul_rn(hi, t);
    const double e = fma(hi, t, -p);
    const double r = (p - rint(p)) + fma(lo, t, e);
    double s, c;
    sincospi(2.0 * r, &s, &c);
    return make_double2(c, -s);
}

__global__ void k_sum_tables(const double* __restrict__ t, long long s0, long long ns, const double2* __restrict__ frow,
                             int nrows, double2* __restrict__ tab) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const int r = blockIdx.y;
    if (r >= nrows) return;
    const double2 F = frow[r];
    tab[(long long)r * ns + s] = cis_turns_exact_dd(F.x, F.y, t[s0 + s]);
}

struct Pre {
    double2 a, pw, d;
    double wt;
};

__global__ void __launch_bounds__(NTHREADS) k_trig_sums(const __grid_constant__ SumArgs a) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int zb = blockIdx.x, prob = blockIdx.y;
    const long long s_begin = a.start0 + (long long)prob * a.hop;
    const int nchunks = (a.n + KC - 1) / KC;
    const int nzb = a.nbm + a.nbp;
    auto load = [&](int c) {
        Pre p;
        const int idx = c * KC + lane;
        const bool valid = idx < a.n && s_begin + idx < a.s_end;
        long long s = s_begin + idx;
        if (!valid) s = min(s_begin + (long long)a.n, a.s_end) - 1;
        const long long si = s - a.tbl_base;
        p.a = a.tab[(long long)zb * a.tbl_ns + si];
        p.pw = a.tab[(long long)(nzb + w) * a.tbl_ns + si];
        p.d = a.del[si];
        double wt = 1.0;
        if (a.W) wt = a.W[a.w_abs ? s : (s - s_begin)];
        p.wt = valid ? wt : 0.0;
        return p;
    };
    double2 acc[GRP];
#pragma unroll
    for (int j = 0; j < GRP; j++) acc[j] = make_double2(0.0, 0.0);
    Pre pn = load(0);
    for (int c = 0; c < nchunks; c++) {
        const Pre p = pn;
        if (c + 1 < nchunks) pn = load(c + 1);
        double2 z = chain_rotate(p.a, p.pw);  // anchor of this warp's group of 8 sums
        z.x *= p.wt;
        z.y *= p.wt;
#pragma unroll
        for (int j = 0; j < GRP; j++) {
            acc[j].x += z.x;
            acc[j].y += z.y;
            z = chain_rotate(z, p.d);
        }
    }
    double2* Zp = a.Z + (long long)prob * a.strideZ + (long long)zb * FB + w * GRP;
#pragma unroll
    for (int j = 0; j < GRP; j++) {
        double vx = acc[j].x, vy = acc[j].y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vx += __shfl_xor_sync(0xffffffffu, vx, o);
            vy += __shfl_xor_sync(0xffffffffu, vy, o);
        }
        if (lane == 0) Zp[j] = make_double2(vx, vy);
    }
}

// out[i] (+)= sum over `nparts` partial Z arrays (sample splits of one problem)
__global__ void k_sum_parts(double2* __restrict__ out, const double2* __restrict__ parts, int count, long long stride,
                            int nparts, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double2 s = accumulate ? out[i] : make_double2(0.0, 0.0);
    for (int p = 0; p < nparts; p++) {
        const double2 v = parts[(long long)p * stride + i];
        s.x += v.x;
        s.y += v.y;
    }
    out[i] = s;
}

__global__ void __launch_bounds__(NTHREADS) k_gram_fill(const __grid_constant__ FillArgs a) {
    int I, J;
    tile_ij(blockIdx.x, I, J);
    const int prob = blockIdx.y;
    const double2* Zm = a.Z + (long long)prob * a.strideZ;
    const double2* Zp = Zm + (long long)a.nbm * FB;
    const int Np = a.nblk * TB;
    double* Gp = a.G + (long long)prob * a.strideG;
    const double h = 0.5 * a.gscale;
    for (int e = threadIdx.x; e < TB * TB; e += NTHREADS) {
        const int r = e >> 7, c = e & (TB - 1);
        const int pi = r >> 6, pj = c >> 6;
        const int i = I * FB + (r & (FB - 1)), j = J * FB + (c & (FB - 1));
        double v = 0.0;
        const bool dummy = i >= a.ncc || j >= a.ncc || (a.zero_first && ((pi == 1 && i == 0) || (pj == 1 && j == 0)));
        if (!dummy) {
            const int md = i - j;
            const double2 zm = __ldg(Zm + (md < 0 ? -md : md));
            const double2 zp = __ldg(Zp + (i + j));
            const double sm = md < 0 ? -zm.y : zm.y;  // Im Z-(i - j)
            if (pi == pj)
                v = pi == 0 ? zm.x + zp.x : zm.x - zp.x;
            else
                v = pi == 1 ? zp.y + sm : zp.y - sm;
            v *= h;
        }
        Gp[(long long)(I * TB + r) * Np + J * TB + c] = v;
    }
}

}  // namespace

void launch_sum_tables(const double* t, long long s0, long long ns, const double2* frow_dev, int nrows, double2* tab,
                       cudaStream_t st) {
    dim3 grid((unsigned)((ns + 255) / 256), nrows);
    k_sum_tables<<<grid, 256, 0, st>>>(t, s0, ns, frow_dev, nrows, tab);
}

int launch_trig_sums(const SumArgs& a, int nproblems, cudaStream_t st) {
    int launched = 0;
    for (int p0 = 0; p0 < nproblems; p0 += 32768) {
        const int np = nproblems - p0 < 32768 ? nproblems - p0 : 32768;
        SumArgs b = a;
        b.start0 = a.start0 + (long long)p0 * a.hop;
        b.Z = a.Z + (long long)p0 * a.strideZ;
        k_trig_sums<<<dim3(a.nbm + a.nbp, np), NTHREADS, 0, st>>>(b);
        launched++;
    }
    return launched;
}

void launch_sum_parts(double2* out, const double2* parts, int count, long long stride, int nparts, int accumulate,
                      cudaStream_t st) {
    k_sum_parts<<<(count + 255) / 256, 256, 0, st>>>(out, parts, count, stride, nparts, accumulate);
}

int launch_gram_fill(const FillArgs& a, int nproblems, cudaStream_t st) {
    int launched = 0;
    const int ntiles = a.nblk * (a.nblk + 1) / 2;
    for (int p0 = 0; p0 < nproblems; p0 += 32768) {
        const int np = nproblems - p0 < 32768 ? nproblems - p0 : 32768;
        FillArgs b = a;
        b.Z = a.Z + (long long)p0 * a.strideZ;
        b.G = a.G + (long long)p0 * a.strideG;
        k_gram_fill<<<dim3(ntiles, np), NTHREADS, 0, st>>>(b);
        launched++;
    }
    return launched;
}

// Frequencies of the table rows as double-double (hi, lo): [0, nbm) -> (64 b) df; [nbm, nbm + nbp) -> 2 f0 + (64 b) df;
// then GRP rows (8 g) df.  Integer x double products and the sum with 2 f0 are carried exactly.
void structured_row_freqs(double f0, double df, int nbm, int nbp, double* out /* 2 * (nbm + nbp + GRP) */) {
    auto two_prod = [](double a, double b, double& lo) {
        const double p = a * b;
        lo = fma(a, b, -p);
        return p;
    };
    auto two_sum = [](double a, double b, double& lo) {
        const double s = a + b, bb = s - a;
        lo = (a - (s - bb)) + (b - bb);
        return s;
    };
    int r = 0;
    for (int b = 0; b < nbm; b++, r++) out[2 * r] = two_prod((double)(FB * b), df, out[2 * r + 1]);
    for (int b = 0; b < nbp; b++, r++) {
        double pl, sl;
        const double ph = two_prod((double)(FB * b), df, pl);
        out[2 * r] = two_sum(2.0 * f0, ph, sl);
        out[2 * r + 1] = sl + pl;
    }
    for (int g = 0; g < GRP; g++, r++) out[2 * r] = two_prod((double)(GRP * g), df, out[2 * r + 1]);
}

}  // namespace lpvs
