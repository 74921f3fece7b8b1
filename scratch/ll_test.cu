// Standalone driver of the left-looking LU determinant kernel (tuning loop without the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 [-DWFO_LL_CLOCKS] -DNTV=13 scratch/ll_test.cu -o scratch/ll_test
#include <vector>
#include <random>
#include <cmath>
#include <algorithm>
#include <numeric>
#include "../pywfo_b200/csrc/det_lu_ll.cuh"
namespace wfo { thread_local char g_err[512]; }
using namespace wfo;
#ifndef NTV
#define NTV 13
#endif
static double host_det(const std::vector<double> &S, int ld, const int *rows, const int *cols, int n) {
    std::vector<long double> a((size_t)n * n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[(size_t)i * n + j] = S[(size_t)rows[i] * ld + cols[j]];
    long double det = 1;
    for (int j = 0; j < n; j++) {
        int p = j;
        for (int i = j + 1; i < n; i++) if (fabsl(a[(size_t)i * n + j]) > fabsl(a[(size_t)p * n + j])) p = i;
        if (p != j) { for (int c = 0; c < n; c++) std::swap(a[(size_t)p * n + c], a[(size_t)j * n + c]); det = -det; }
        long double piv = a[(size_t)j * n + j];
        det *= piv;
        if (piv == 0) return 0;
        for (int i = j + 1; i < n; i++) {
            long double l = a[(size_t)i * n + j] / piv;
            for (int c = j + 1; c < n; c++) a[(size_t)i * n + c] -= l * a[(size_t)j * n + c];
        }
    }
    return (double)det;
}
int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 8 * NTV - 4;
    const int mode = argc > 2 ? atoi(argv[2]) : 0;   // 0 near-identity excitation lists, 1 random
    const int kb = argc > 3 ? atoi(argv[3]) : 256, kk = argc > 4 ? atoi(argv[4]) : 512;
    const int n_mo = 5 * n;
    std::mt19937_64 rng(1234 + n);
    std::normal_distribution<double> nd;
    std::vector<double> S((size_t)n_mo * n_mo);
    for (int i = 0; i < n_mo; i++) for (int j = 0; j < n_mo; j++) S[(size_t)i * n_mo + j] = mode ? nd(rng) : ((i == j ? 1.0 : 0.0) + 0.05 * nd(rng));
    std::vector<int> rows((size_t)kb * n), cols((size_t)kk * n);
    auto fill = [&](std::vector<int> &v, int cnt) {
        for (int k = 0; k < cnt; k++) {
            if (mode) {
                std::vector<int> perm(n_mo); std::iota(perm.begin(), perm.end(), 0); std::shuffle(perm.begin(), perm.end(), rng);
                for (int i = 0; i < n; i++) v[(size_t)k * n + i] = perm[i];
            } else {
                int f = rng() % n, t = rng() % (n_mo - n), c = 0;
                for (int i = 0; i < n; i++) if (i != f) v[(size_t)k * n + c++] = i;
                v[(size_t)k * n + c] = n + t;
            }
        }
    };
    fill(rows, kb); fill(cols, kk);
    double *dS, *dD; int *dr, *dc;
    cudaMalloc(&dS, S.size() * 8); cudaMalloc(&dD, (size_t)kb * kk * 8); cudaMalloc(&dr, rows.size() * 4); cudaMalloc(&dc, cols.size() * 4);
    cudaMemcpy(dS, S.data(), S.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dr, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice);
    #ifndef MGV
#define MGV false
#endif
    auto kern = det_table_ll_kernel<NTV, MGV>;
    constexpr int LL_WARPS = LLWarps<NTV, MGV>::W;
    const int smem = LL_WARPS * LLCfg<NTV, MGV>::BYTES;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LL_WARPS * 32, smem);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * per_sm;
    double *dG = nullptr;
    if (MGV) cudaMalloc(&dG, (size_t)grid * LL_WARPS * ll_scratch_doubles(NTV) * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, sum = 0; const int reps = 10;
    for (int r = 0; r < reps + 5; r++) {
#ifdef WFO_LL_CLOCKS
        long long z[8] = {0}; cudaMemcpyToSymbol(g_ll_clk, z, sizeof(z));
#endif
        cudaEventRecord(e0);
        kern<<<grid, LL_WARPS * 32, smem>>>(dS, n_mo, n, dr, kb, dc, kk, dD, dG);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 5) { best = std::min(best, ms); sum += ms; }
    }
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    std::vector<double> D((size_t)kb * kk);
    cudaMemcpy(D.data(), dD, D.size() * 8, cudaMemcpyDeviceToHost);
    double worst = 0; int bad = 0;
    for (int k = 0; k < kb; k += 37) for (int l = 0; l < kk; l += 61) {
        const double ref = host_det(S, n_mo, &rows[(size_t)k * n], &cols[(size_t)l * n], n);
        const double e = fabs(D[(size_t)k * kk + l] - ref) / fabs(ref);
        worst = std::max(worst, e);
        if (!(e < 1e-9)) bad++;
    }
    const double fdet = n * (n - 1) / 2.0 + n * (n - 1) * (2.0 * n - 1) / 3.0 + n - 1;
    const double ms = sum / reps, tf = kb * (double)kk * fdet / ms / 1e9;
    const double clk_mat = ms * 1e-3 * 1.965e9 / (kb * (double)kk / (grid * LL_WARPS));
    printf("MG=%d n=%d NT=%d mode=%d pairs=%d warps/SM=%d ms=%.3f (best %.3f) TF/s=%.2f frac=%.3f clk/matrix/warp=%.0f max_rel_err=%.2e bad=%d\n", (int)MGV, n, NTV, mode,
           kb * kk, per_sm * LL_WARPS, ms, best, tf, tf / 37.17, clk_mat, worst, bad);
#ifdef WFO_LL_CLOCKS
    long long c[8]; cudaMemcpyFromSymbol(c, g_ll_clk, sizeof(c));
    const char *names[8] = {"A prefetch issue", "left-looking update", "staging write", "staging read", "8 column steps", "M compute+write", "moved rows", ""};
    double tot = 0; for (int i = 0; i < 7; i++) tot += c[i];
    for (int i = 0; i < 7; i++) printf("   %-20s %5.1f %%  %8.0f clk/matrix\n", names[i], 100.0 * c[i] / tot, c[i] / ((double)kb * kk));
#endif
    return bad ? 2 : 0;
}
