// Cross-process P2P probe (one process per GPU, CUDA IPC over NVLink): can a persistent kernel on GPU a signal a
// persistent kernel on GPU b through peer memory, and what does one hop cost?  Feeds the sharded-ADMM decision.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_probe tools/p2p_probe.cu
// Run on a box with >= 2 GPUs: tools/p2p_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[%d] CUDA %s at %s:%d\n", g_rank, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
static int g_rank = 0;

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
    asm volatile("st.global.release.sys.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.global.acquire.sys.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// flag ping-pong, one thread
__global__ void k_pingpong(long long* local_flag, long long* peer_flag, int iters, int rank, long long* timeout) {
    for (int it = 1; it <= iters; it++) {
        if (rank == 0) st_release_sys(peer_flag, it);
        long long spins = 0;
        while (ld_acquire_sys(local_flag) < it) {
            if (++spins > 200000000LL) { *timeout = it; return; }
        }
        if (rank == 1) st_release_sys(peer_flag, it);
    }
}

// payload exchange: every rank writes `n` doubles into the peer's slot, fences, raises the flag; waits for the peer's
__global__ void k_exchange(double* local_buf, double* peer_buf, long long* local_flag, long long* peer_flag, int n,
                           int iters, long long* timeout) {
    for (int it = 1; it <= iters; it++) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) peer_buf[i] = (double)it + i;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            st_release_sys(peer_flag, it);
            long long spins = 0;
            while (ld_acquire_sys(local_flag) < it) {
                if (++spins > 200000000LL) { *timeout = it; break; }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && local_buf[n - 1] != (double)it + (n - 1)) *timeout = -it;
    }
}

int main() {
    int p01[2], p10[2];
    if (pipe(p01) || pipe(p10)) return 1;
    pid_t pid = fork();  // before ANY CUDA call: a forked child cannot reuse the parent's CUDA state
    g_rank = pid == 0 ? 1 : 0;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 2) { printf("needs 2 GPUs (found %d)\n", ndev); return 0; }
    int rfd = g_rank == 0 ? p10[0] : p01[0], wfd = g_rank == 0 ? p01[1] : p10[1];
    CK(cudaSetDevice(g_rank));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, g_rank, 1 - g_rank));
    char* buf;
    CK(cudaMalloc(&buf, 1 << 20));
    CK(cudaMemset(buf, 0, 1 << 20));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mine, theirs;
    CK(cudaIpcGetMemHandle(&mine, buf));
    if (write(wfd, &mine, sizeof(mine)) != sizeof(mine)) return 1;
    if (read(rfd, &theirs, sizeof(theirs)) != sizeof(theirs)) return 1;
    char* peer;
    CK(cudaIpcOpenMemHandle((void**)&peer, theirs, cudaIpcMemLazyEnablePeerAccess));
    // layout: [0,8) flag A, [64,72) flag B, [128,136) timeout, [4096, ...) payload
    long long* lflag = (long long*)buf; long long* pflag = (long long*)peer;
    long long* lflag2 = (long long*)(buf + 64); long long* pflag2 = (long long*)(peer + 64);
    long long* tmo = (long long*)(buf + 128);
    double* lpay = (double*)(buf + 4096); double* ppay = (double*)(peer + 4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    char tok = 1;  // crude rendezvous so both kernels start close together
    if (write(wfd, &tok, 1) != 1 || read(rfd, &tok, 1) != 1) return 1;
    const int iters = 20000;
    cudaEventRecord(e0);
    k_pingpong<<<1, 1>>>(lflag, pflag, iters, g_rank, tmo);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long t = 0; CK(cudaMemcpy(&t, tmo, 8, cudaMemcpyDeviceToHost));
    printf("[%d] canAccessPeer=%d  flag ping-pong: %.2f us per round trip (%.2f us one way), timeout=%lld\n", g_rank, can,
           ms * 1e3 / iters, ms * 1e3 / iters / 2, t);
    for (int n : {16, 2048, 16384}) {
        if (write(wfd, &tok, 1) != 1 || read(rfd, &tok, 1) != 1) return 1;
        CK(cudaMemset(buf + 64, 0, 8));
        CK(cudaDeviceSynchronize());
        if (write(wfd, &tok, 1) != 1 || read(rfd, &tok, 1) != 1) return 1;
        const int it2 = 5000;
        cudaEventRecord(e0);
        k_exchange<<<1, 512>>>(lpay, ppay, lflag2, pflag2, n, it2, tmo);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(&t, tmo, 8, cudaMemcpyDeviceToHost));
        printf("[%d] exchange of %6d doubles (%6.1f KB) + flag, one CTA: %.2f us per iteration, timeout/err=%lld\n", g_rank, n,
               n * 8 / 1024.0, ms * 1e3 / it2, t);
    }
    CK(cudaIpcCloseMemHandle(peer));
    if (g_rank == 0) { int st; wait(&st); }
    return 0;
}
