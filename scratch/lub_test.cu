// Standalone timing of base_inverse_kernel (scratch): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 scratch/lub_test.cu -o scratch/lub_test
//   lub_test <mode>        table over n (warm / L2-flushed launches; mode 0 near-identity, 1 random)
//   lub_test <mode> <n>    200 inversions inside ONE launch (per-inversion time without launch overhead; ncu sampling)
// Round 2 went through four designs with it (us at n = 100): round 1's one block barrier per column 81; panels of 8
// columns published per PANEL through mbarriers 97-100 (the next owner's apply sits on the critical path); published per
// COLUMN 82; the same with predicated stores instead of branches and a software-pipelined pivot step 68 (shipped).
// A lone warp issues one dependent instruction every ~5 cycles, a pivot step is ~140 of them: that is the floor here.
#include <vector>
#include <random>
#include <cstdio>
#include <cmath>
#include "../pywfo_b200/csrc/lu_base.cuh"
namespace wfo { thread_local char g_err[512]; }
using namespace wfo;
__global__ void flush_kernel(char *p, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 1; }
__global__ void __launch_bounds__(LUB_THREADS) repeat_kernel(const double *A, int n, double *Ai, BaseInfo *info, int reps) {
    extern __shared__ __align__(16) double lub_tile[];
    double det = 0, minp = 0, maxp = 0; int sing = 0;
    for (int it = 0; it < reps; it++) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < n * n; idx += LUB_THREADS) { const int r = idx / n, c = idx - r * n; lub_tile[r * LUB_LDS + c] = A[(size_t)r * n + c]; }
        __syncthreads();
        lub_invert_tile(lub_tile, n, det, minp, maxp, sing);
    }
    for (int idx = threadIdx.x; idx < n * n; idx += LUB_THREADS) { const int r = idx / n, c = idx - r * n; Ai[(size_t)r * n + c] = lub_tile[r * LUB_LDS + c]; }
    if (threadIdx.x == 0) { info->det = det; info->singular = sing; }
}
int main(int argc, char **argv) {
    if (argc > 2) {   // lub_test <mode> <n>: one launch of the repeat kernel (for ncu sampling)
        const int n = atoi(argv[2]);
        std::mt19937_64 rng(7 + n); std::normal_distribution<double> nd;
        std::vector<double> A((size_t)n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) A[i * n + j] = ((i == j) + 0.05 * nd(rng));
        double *dA, *dAi; BaseInfo *dinfo;
        cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dAi, A.size() * 8); cudaMalloc(&dinfo, sizeof(BaseInfo));
        cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(repeat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lub_smem_bytes(128));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        repeat_kernel<<<1, LUB_THREADS, lub_smem_bytes(n)>>>(dA, n, dAi, dinfo, 200);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("n=%d: %.2f us per inversion (200 in one launch): %s\n", n, ms * 1e3 / 200, cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    char *fl; const size_t FB = 256u << 20; cudaMalloc(&fl, FB);
    cudaFuncSetAttribute(base_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lub_smem_bytes(128));
    for (int n : {5, 16, 34, 64, 100, 128}) {
        std::mt19937_64 rng(7 + n); std::normal_distribution<double> nd;
        std::vector<double> A((size_t)n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) A[i * n + j] = mode ? nd(rng) : ((i == j) + 0.05 * nd(rng));
        double *dA, *dAi; BaseInfo *dinfo;
        cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dAi, A.size() * 8); cudaMalloc(&dinfo, sizeof(BaseInfo));
        cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float warm = 0, cold = 0;
        for (int r = 0; r < 8; r++) {
            cudaEventRecord(e0);
            base_inverse_kernel<<<1, LUB_THREADS, lub_smem_bytes(n)>>>(dA, n, n, dAi, n, dinfo, 1, 0, nullptr, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 3) warm += ms / 5;
        }
        for (int r = 0; r < 5; r++) {
            flush_kernel<<<296, 1024>>>(fl, FB); cudaDeviceSynchronize();
            cudaEventRecord(e0);
            base_inverse_kernel<<<1, LUB_THREADS, lub_smem_bytes(n)>>>(dA, n, n, dAi, n, dinfo, 1, 0, nullptr, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); cold += ms / 5;
        }
        std::vector<double> Ai(A.size()); cudaMemcpy(Ai.data(), dAi, Ai.size() * 8, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = 0; for (int k = 0; k < n; k++) s += A[i * n + k] * Ai[k * n + j]; err = fmax(err, fabs(s - (i == j))); }
        printf("n=%3d warm %.1f us  cold(L2 flushed) %.1f us  |A Ai - I| %.1e\n", n, warm * 1e3, cold * 1e3, err);
#ifdef WFO_LUB_CLOCKS
        long long c[16]; cudaMemcpyFromSymbol(c, g_lub_clk, sizeof(c));
        const char *nm[10] = {"load regs", "first publish", "(loop top)", "row exchange", "mcol read", "lookahead", "bulk update", "barrier", "(end loop)", "write-out"};
        for (int i = 0; i < 10; i++) printf("    %-14s %8lld clk\n", nm[i], c[i]);
#endif
        cudaFree(dA); cudaFree(dAi); cudaFree(dinfo);
    }
    return 0;
}
