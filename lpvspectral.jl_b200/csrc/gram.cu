// Fused basis synthesis -> DMMA SYRK.  See gram.cuh for the layout contract.
//
// Per CTA: one 128x128 tile (I,J), I >= J, of one problem's Gram; 8 warps as 4(M) x 2(N), warp tile 32x64,
// 64 FP64 accumulators per thread.  The k-dimension (samples) is consumed in chunks of 32: while the DMMAs run on
// the current chunk's smem tiles, the same threads synthesise the next chunk's tiles into the other buffer
// (software pipelined inside the k4-step loop because DMMA and DFMA share one pipe -- tools/fp64_probe.cu).
//
// GRAM_CHAIN synthesis (uniform frequency grid): thread (warp w, lane l) owns sample l of the chunk and the 8
// frequencies 64*blk+8w .. +7; it starts from an exact anchor e^{-i 2 pi f t} (table, double-double phase) and
// advances with the per-sample rotation e^{-i 2 pi df t}: 4 FP64 ops per (cos,-sin) pair instead of ~35.
#include "gram.cuh"

namespace lpvs {

namespace {

constexpr int STAGE_D = 2 * TILE_D + 8 * LDT;  // I tile, J tile, rhs tile (8 rows)

struct Pref {
    double2 aI, aJ, d;  // chain: anchors + step rotation
    double tt;          // direct: sample position
    long long si;       // table sample index
    double wt, yv;
};

template <int MODE, bool DIAG>
__device__ __forceinline__ Pref load_pref(const GramArgs& a, int c, long long s_begin, int lane, int w, int I,
                                          int J) {
    Pref p{};
    int idx = c * KC + lane;
    bool valid = idx < a.n && s_begin + idx < a.s_end;
    long long s = s_begin + idx;
    if (!valid) s = min(s_begin + (long long)a.n, a.s_end) - 1;
    int idc = (int)(s - s_begin);
    p.si = s - a.tbl_base;
    double wt = 1.0;
    if (a.W) wt = a.W[a.w_abs ? s : (long long)idc];
    p.wt = valid ? wt : 0.0;
    p.yv = 0.0;
    if (DIAG && w < a.nrhs) p.yv = (w == 0 ? a.y : a.u)[s] * p.wt;
    if (MODE == GRAM_CHAIN) {
        p.aI = a.anc[(long long)(I * (FB / GRP) + w) * a.tbl_ns + p.si];
        if (!DIAG) p.aJ = a.anc[(long long)(J * (FB / GRP) + w) * a.tbl_ns + p.si];
        p.d = a.del[p.si];
    } else if (MODE == GRAM_DIRECT) {
        p.tt = a.t[s];
    }
    return p;
}

// value of complex column cc at the prefetched sample (slow paths)
template <int MODE>
__device__ __forceinline__ double2 synth_elem(const GramArgs& a, const Pref& p, int cc) {
    if (MODE == GRAM_DIRECT) {
        return cis_reference(a.f[cc], p.tt);
    } else {
        int fi = cc % a.lpv_nf, ki = cc / a.lpv_nf;
        double2 e = a.E[(long long)fi * a.tbl_ns + p.si];
        double k = a.Kt[(long long)ki * a.tbl_ns + p.si];
        return make_double2(e.x * k, e.y * k);
    }
}

template <int MODE, bool DIAG>
__device__ __forceinline__ void gram_tile(const GramArgs& a, int I, int J, int prob, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wm = w & 3, wn = w >> 2;
    const long long s_begin = a.start0 + (long long)prob * a.hop;
    const int nchunks = (a.n + KC - 1) / KC;

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    double accb[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) accb[i][0] = accb[i][1] = 0.0;

    // zero the rhs tiles once (rows >= nrhs stay zero)
    for (int q = tid; q < 8 * LDT; q += NTHREADS) {
        smem[2 * TILE_D + q] = 0.0;
        smem[STAGE_D + 2 * TILE_D + q] = 0.0;
    }

    const int ccI0 = I * FB + w * GRP, ccJ0 = J * FB + w * GRP;
    const int rowc = (w * GRP) * LDT + lane;         // smem offset of the real-part row for j = 0
    const int rows = (FB + w * GRP) * LDT + lane;    // second-part row

    // synthesise element j (of 8) of a chunk into stage buffer `st`
    auto synth_step = [&](const Pref& p, double2& zI, double2& zJ, int j, double* st) {
        double2 vI, vJ;
        if (MODE == GRAM_CHAIN) {
            vI = zI;
            vJ = DIAG ? zI : zJ;
        } else {
            vI = (ccI0 + j < a.ncc) ? synth_elem<MODE>(a, p, ccI0 + j) : make_double2(0.0, 0.0);
            vJ = DIAG ? vI : ((ccJ0 + j < a.ncc) ? synth_elem<MODE>(a, p, ccJ0 + j) : make_double2(0.0, 0.0));
        }
        bool okI = ccI0 + j < a.ncc, okJ = ccJ0 + j < a.ncc;
        double* sI = st;
        double* sJ = st + TILE_D;
        sI[rowc + j * LDT] = okI ? vI.x : 0.0;
        sI[rows + j * LDT] = okI ? vI.y : 0.0;
        sJ[rowc + j * LDT] = okJ ? vJ.x * p.wt : 0.0;
        sJ[rows + j * LDT] = okJ ? vJ.y * p.wt : 0.0;
        if (MODE == GRAM_CHAIN) {
            double nx = zI.x * p.d.x - zI.y * p.d.y;
            double ny = zI.x * p.d.y + zI.y * p.d.x;
            zI = make_double2(nx, ny);
            if (!DIAG) {
                nx = zJ.x * p.d.x - zJ.y * p.d.y;
                ny = zJ.x * p.d.y + zJ.y * p.d.x;
                zJ = make_double2(nx, ny);
            }
        }
    };
    auto synth_rhs = [&](const Pref& p, double* st) {
        if (DIAG && w < a.nrhs) st[2 * TILE_D + w * LDT + lane] = p.yv;
    };

    // prologue: chunk 0 into stage 0
    Pref p1 = load_pref<MODE, DIAG>(a, 0, s_begin, lane, w, I, J);
    {
        double2 zI = p1.aI, zJ = p1.aJ;
#pragma unroll
        for (int j = 0; j < GRP; j++) synth_step(p1, zI, zJ, j, smem);
        synth_rhs(p1, smem);
    }
    if (nchunks > 1) p1 = load_pref<MODE, DIAG>(a, 1, s_begin, lane, w, I, J);
    __syncthreads();

    const int fragA = (32 * wm + (lane >> 2)) * LDT + (lane & 3);
    const int fragB = (64 * wn + (lane >> 2)) * LDT + (lane & 3);
    const int fragY = (lane >> 2) * LDT + (lane & 3);

    for (int c = 0; c < nchunks; c++) {
        double* cur = smem + (c & 1) * STAGE_D;
        double* nxt = smem + ((c & 1) ^ 1) * STAGE_D;
        const bool have_next = (c + 1 < nchunks);
        Pref p2 = p1;
        if (c + 2 < nchunks) p2 = load_pref<MODE, DIAG>(a, c + 2, s_begin, lane, w, I, J);
        double2 zI = p1.aI, zJ = p1.aJ;
        const double* pa = cur + fragA;
        const double* pb = cur + TILE_D + fragB;
#pragma unroll
        for (int kk = 0; kk < KC / 4; kk++) {
            if (have_next) synth_step(p1, zI, zJ, kk, nxt);
            mma_step(pa, pb, kk, acc);
            if (DIAG && wn == 0) {
                double by = cur[2 * TILE_D + fragY + 4 * kk];
#pragma unroll
                for (int i = 0; i < 4; i++) dmma884(accb[i][0], accb[i][1], pa[i * 8 * LDT + 4 * kk], by);
            }
        }
        if (have_next) synth_rhs(p1, nxt);
        p1 = p2;
        __syncthreads();
    }

    // epilogue
    const int Np = a.nblk * TB;
    double* Gp = a.G + (long long)prob * a.strideG;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int row = I * TB + 32 * wm + 8 * i + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            int col = J * TB + 64 * wn + 8 * j + 2 * (lane & 3);
            double2 v = make_double2(acc[i][j][0] * a.gscale, acc[i][j][1] * a.gscale);
            *reinterpret_cast<double2*>(Gp + (long long)row * Np + col) = v;
        }
    }
    if (DIAG && wn == 0 && a.B) {
        double* Bp = a.B + (long long)prob * a.strideB;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int row = I * TB + 32 * wm + 8 * i + (lane >> 2);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                int r = 2 * (lane & 3) + e;
                if (r < a.nrhs) Bp[(long long)r * Np + row] = accb[i][e] * a.bscale;
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) k_gram(const __grid_constant__ GramArgs a) {
    extern __shared__ __align__(16) double smem[];
    int I, J;
    tile_ij(blockIdx.x, I, J);
    if (I == J)
        gram_tile<MODE, true>(a, I, J, blockIdx.y, smem);
    else
        gram_tile<MODE, false>(a, I, J, blockIdx.y, smem);
}

__global__ void k_anchor_table(const double* __restrict__ t, long long s0, long long ns,
                               const double* __restrict__ f, int Nf, int ngroups, double f0, double df,
                               double2* __restrict__ anc, double2* __restrict__ del) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= ns) return;
    int g = blockIdx.y;
    double tt = t[s0 + s];
    if (g == ngroups) {
        del[s] = cis_turns_exact(df, tt);
    } else {
        int k = g * GRP;
        double fk = (k < Nf) ? f[k] : fma((double)k, df, f0);
        anc[(long long)g * ns + s] = cis_turns_exact(fk, tt);
    }
}

__global__ void k_lpv_tables(const double* __restrict__ X, const double* __restrict__ V, long long N,
                             const double* __restrict__ w, int Nf, int Nvv, const double* __restrict__ centers,
                             double gamma, int coulomb, int normalize, double2* __restrict__ E,
                             double* __restrict__ Kt) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    double x = X[s], v = V[s];
    // E: exp(-i * fl(w_f * X_n))  (src/lsfft.jl:244: exp.(im*w.*X) then the trailing adjoint conjugates)
    for (int fi = 0; fi < Nf; fi++) E[(long long)fi * N + s] = cis_of_phase(__dmul_rn(w[fi], x));
    // K: exp(-gamma (v - c_k)^2) [* (sign v == sign c_k)] [/ sum]   (src/lsfft.jl:195-207)
    double sum = 0.0;
    double sv = (v > 0.0) - (v < 0.0);
    for (int k = 0; k < Nvv; k++) {
        double c = centers[k];
        double dv = v - c;
        double val = exp(-gamma * (dv * dv));
        if (coulomb) {
            double sc = (c > 0.0) - (c < 0.0);
            if (sv != sc) val = 0.0;
        }
        sum += val;
        Kt[(long long)k * N + s] = val;
    }
    if (normalize) {
        for (int k = 0; k < Nvv; k++) Kt[(long long)k * N + s] /= sum;
    }
}

}  // namespace

size_t gram_smem_bytes() { return 2 * STAGE_D * sizeof(double); }

void launch_gram(int mode, const GramArgs& a, int nproblems, cudaStream_t st) {
    static bool attr_done = false;
    size_t smem = gram_smem_bytes();
    if (!attr_done) {
        cudaFuncSetAttribute(k_gram<GRAM_CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_gram<GRAM_DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_gram<GRAM_LPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    int ntiles = a.nblk * (a.nblk + 1) / 2;
    // gridDim.y is limited to 65535: launch in slabs
    for (int p0 = 0; p0 < nproblems; p0 += 32768) {
        int np = nproblems - p0 < 32768 ? nproblems - p0 : 32768;
        GramArgs b = a;
        b.start0 = a.start0 + (long long)p0 * a.hop;
        b.G = a.G + (long long)p0 * a.strideG;
        if (a.B) b.B = a.B + (long long)p0 * a.strideB;
        dim3 grid(ntiles, np);
        if (mode == GRAM_CHAIN)
            k_gram<GRAM_CHAIN><<<grid, NTHREADS, smem, st>>>(b);
        else if (mode == GRAM_DIRECT)
            k_gram<GRAM_DIRECT><<<grid, NTHREADS, smem, st>>>(b);
        else
            k_gram<GRAM_LPV><<<grid, NTHREADS, smem, st>>>(b);
    }
}

void launch_anchor_table(const double* t, long long s0, long long ns, const double* f, int Nf, int ngroups,
                         double f0, double df, double2* anc, double2* del, cudaStream_t st) {
    dim3 grid((unsigned)((ns + 255) / 256), ngroups + 1);
    k_anchor_table<<<grid, 256, 0, st>>>(t, s0, ns, f, Nf, ngroups, f0, df, anc, del);
}

void launch_lpv_tables(const double* X, const double* V, long long N, const double* w, int Nf, int Nvv,
                       const double* centers, double gamma, int coulomb, int normalize, double2* E, double* Kt,
                       cudaStream_t st) {
    k_lpv_tables<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(X, V, N, w, Nf, Nvv, centers, gamma, coulomb,
                                                              normalize, E, Kt);
}

}  // namespace lpvs
