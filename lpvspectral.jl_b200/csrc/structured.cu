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
// The right-hand sides are sums of the same kind, Zy(k) = sum_s W_s y_s e^{-i 2 pi (f0 + k df) t_s}: A'Wy = (Re Zy, Im Zy).  With
// f0 = 0 (the reference's default grids) Z- is the head of Z+, so a problem costs 2 Nf sums for G and Nf per right-hand side.
//
// k_sum_tables     per-sample table rows e^{-i 2 pi F_r t_s}: block anchors of every sum family, the 8 group powers, the step
// k_trig_sums      one CTA per (64 sums, problem): warp = 8 consecutive m (angle-addition chain), lane = sample of a chunk
// k_gram_fill      one CTA per (lower 128 x 128 tile, problem): the identities above, internal column layout of gram.cuh
// k_rhs_from_sums  b in the internal layout
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

// table row r < nzb: block anchor r; rows nzb .. nzb + GRP: group powers and the step, which sit behind the block rows of the
// FULL layout (two right-hand sides) in frow
__global__ void k_sum_tables(const double* __restrict__ t, long long s0, long long ns, const double2* __restrict__ frow,
                             int nzb, int nzb_full, double2* __restrict__ tab) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    const int r = blockIdx.y;
    const double2 F = frow[r < nzb ? r : r - nzb + nzb_full];
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
    // blocks [0, nzg) are sums for G (weight W), then nby blocks per right-hand side (weight W y, W u)
    const double* ysrc = zb < a.nzg ? nullptr : (zb < a.nzg + a.nby ? a.y : a.u);
    auto load = [&](int c) {
        Pre p;
        const int idx = c * KC + lane;
        const bool valid = idx < a.n && s_begin + idx < a.s_end;
        long long s = s_begin + idx;
        if (!valid) s = min(s_begin + (long long)a.n, a.s_end) - 1;
        const long long si = s - a.tbl_base;
        p.a = a.tab[(long long)zb * a.tbl_ns + si];
        p.pw = a.tab[(long long)(a.nzb + w) * a.tbl_ns + si];
        p.d = a.tab[(long long)(a.nzb + GRP) * a.tbl_ns + si];
        double wt = 1.0;
        if (a.W) wt = a.W[a.w_abs ? s : (s - s_begin)];
        if (ysrc) wt *= ysrc[s];
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
    const double2* Zm = a.Z + (long long)prob * a.strideZ + a.zm_off;
    const double2* Zp = a.Z + (long long)prob * a.strideZ + a.zp_off;
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

// b = A' diag(W) [y u] from the Zy / Zu sums: internal layout, dummies zero
__global__ void k_rhs_from_sums(const __grid_constant__ FillArgs a, int nrhs, double bscale, double* __restrict__ B,
                                long long strideB) {
    const int Np = a.nblk * TB;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Np) return;
    const int prob = blockIdx.y;
    const int part = (p >> 6) & 1, k = (p >> 7) * FB + (p & (FB - 1));
    const bool dummy = k >= a.ncc || (a.zero_first && part == 1 && k == 0);
    for (int r = 0; r < nrhs; r++) {
        double v = 0.0;
        if (!dummy) {
            const double2 z = a.Z[(long long)prob * a.strideZ + a.zy_off[r] + k];
            v = bscale * (part == 0 ? z.x : z.y);
        }
        B[(long long)prob * strideB + (long long)r * Np + p] = v;
    }
}

}  // namespace

void launch_rhs_from_sums(const FillArgs& a, int nrhs, double bscale, double* B, long long strideB, int nproblems,
                          cudaStream_t st) {
    const int Np = a.nblk * TB;
    for (int p0 = 0; p0 < nproblems; p0 += 32768) {
        const int np = nproblems - p0 < 32768 ? nproblems - p0 : 32768;
        FillArgs b = a;
        b.Z = a.Z + (long long)p0 * a.strideZ;
        k_rhs_from_sums<<<dim3((Np + 255) / 256, np), 256, 0, st>>>(b, nrhs, bscale, B + (long long)p0 * strideB, strideB);
    }
}

void launch_sum_tables(const double* t, long long s0, long long ns, const double2* frow_dev, int nzb, int nzb_full,
                       double2* tab, cudaStream_t st) {
    dim3 grid((unsigned)((ns + 255) / 256), nzb + GRP + 1);
    k_sum_tables<<<grid, 256, 0, st>>>(t, s0, ns, frow_dev, nzb, nzb_full, tab);
}

int launch_trig_sums(const SumArgs& a, int nproblems, cudaStream_t st) {
    int launched = 0;
    for (int p0 = 0; p0 < nproblems; p0 += 32768) {
        const int np = nproblems - p0 < 32768 ? nproblems - p0 : 32768;
        SumArgs b = a;
        b.start0 = a.start0 + (long long)p0 * a.hop;
        b.Z = a.Z + (long long)p0 * a.strideZ;
        k_trig_sums<<<dim3(a.nzb, np), NTHREADS, 0, st>>>(b);
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

// Layout of the sums of one problem for a grid f0 + k df with nrhs right-hand sides, and the double-double (hi, lo) frequencies
// of the table rows: one row per 64-sum block (first frequency of the block), then GRP rows (8 g) df, then the step df.
// Integer x double products and the sums with f0 / 2 f0 are carried exactly.
StructuredLayout structured_layout(double f0, int Nf, int nrhs) {
    StructuredLayout L{};
    const int nbm = (Nf + FB - 1) / FB, nbp = (2 * Nf - 1 + FB - 1) / FB;
    L.alias = f0 == 0.0;  // Z-(m) = Z+(m): one family serves both
    L.zm_off = 0;
    L.zp_off = L.alias ? 0 : nbm * FB;
    L.nzg = L.alias ? nbp : nbm + nbp;
    L.nby = nrhs > 0 ? nbm : 0;
    L.zy_off[0] = L.nzg * FB;
    L.zy_off[1] = (L.nzg + L.nby) * FB;
    L.nzb = L.nzg + (nrhs > 0 ? nrhs : 0) * L.nby;
    L.nrows = L.nzb + GRP + 1;
    return L;
}

void structured_row_freqs(double f0, double df, int Nf, int nrhs, double* out /* 2 * nrows */) {
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
    const StructuredLayout L = structured_layout(f0, Nf, nrhs);
    const int nbm = (Nf + FB - 1) / FB, nbp = (2 * Nf - 1 + FB - 1) / FB;
    int r = 0;
    auto family = [&](double base, int nblocks) {
        for (int b = 0; b < nblocks; b++, r++) {
            double pl, sl;
            const double ph = two_prod((double)(FB * b), df, pl);
            out[2 * r] = two_sum(base, ph, sl);
            out[2 * r + 1] = sl + pl;
        }
    };
    if (!L.alias) family(0.0, nbm);
    family(2.0 * f0, nbp);
    for (int q = 0; q < (nrhs > 0 ? nrhs : 0); q++) family(f0, nbm);
    for (int g = 0; g < GRP; g++, r++) out[2 * r] = two_prod((double)(GRP * g), df, out[2 * r + 1]);
    out[2 * r] = df;
    out[2 * r + 1] = 0.0;
}

}  // namespace lpvs
