// Stand-alone probe #2 for the remaining tcgen05 building blocks of the NRC kernels (run on the B200 via
// tests/test_gpu_tc05_probe.py):
//   T1  SS-mode: A (M=128, K-major, no swizzle) and B (K-major) both in shared memory      -> layer-0 forward
//   T2  TS-mode: A in TMEM, B in shared memory MN-major (row-major [k][n] weights, no transpose) -> backward
//   T3  SS-mode, M=64: A and B both MN-major, K = 128 samples                                -> weight gradients
// MN-major canonical no-swizzle layout (cute/atom/mma_traits_sm100.hpp:171): element (mn, k) lives at
//   (mn/8)*SBO + (mn%8)*2 + (k%8)*16 + (k/8)*LBO   bytes.
// M=64 accumulators occupy lanes (m%16) + 32*(m/16) (mma_traits_sm100.hpp:504-516).
#include "../../nrc_hpm_renderer_b200/csrc/tc05.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

using namespace tc05;

__host__ __device__ inline uint32_t mnmajor_offset(uint32_t mn, uint32_t k, uint32_t lbo, uint32_t sbo) {
    return (mn >> 3) * sbo + (mn & 7u) * 2u + (k & 7u) * 16u + (k >> 3) * lbo;
}

// mode 1: T1, mode 2: T2, mode 3: T3.  A: [M][K] row-major (mode 1,2) or [K][M] row-major (mode 3).
// B: [N][K] row-major (mode 1) or [K][N] row-major (mode 2,3).  D: raw TMEM dump [128 lanes][N].
template <int MODE, int M, int N, int K>
__global__ void __launch_bounds__(128, 1) probe2(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + 32768;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tmem_alloc(&tmem_base_s, 256); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t d_col = 0, a_col = 128;
    // zero the accumulator region so unused lanes read back as 0
    {
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 16; j++) z[j] = 0;
#pragma unroll
        for (int j = 0; j < N; j += 16) tmem_st16(tbase + lane_base + d_col + j, z);
        wait_st();
    }
    if (MODE == 1) {
        for (int i = tid; i < M * K; i += 128) { int m = i / K, k = i % K; *reinterpret_cast<__half*>(a_s + kmajor_offset(m, k, K)) = A[i]; }
        for (int i = tid; i < N * K; i += 128) { int n = i / K, k = i % K; *reinterpret_cast<__half*>(b_s + kmajor_offset(n, k, K)) = B[i]; }
    } else if (MODE == 2) {
        uint32_t a[K / 2];
#pragma unroll
        for (int j = 0; j < K / 2; j++) a[j] = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K)[j];
#pragma unroll
        for (int j = 0; j < K / 2; j += 8) tmem_st8(tbase + lane_base + a_col + j, a + j);
        wait_st();
        for (int i = tid; i < K * N; i += 128) { int k = i / N, n = i % N; *reinterpret_cast<__half*>(b_s + mnmajor_offset(n, k, 128, (K / 8) * 128)) = B[i]; }
    } else {
        for (int i = tid; i < K * M; i += 128) { int k = i / M, m = i % M; *reinterpret_cast<__half*>(a_s + mnmajor_offset(m, k, 128, (K / 8) * 128)) = A[i]; }
        for (int i = tid; i < K * N; i += 128) { int k = i / N, n = i % N; *reinterpret_cast<__half*>(b_s + mnmajor_offset(n, k, 128, (K / 8) * 128)) = B[i]; }
    }
    fence_proxy_async_smem();
    fence_before();
    __syncthreads();
    if (tid == 0) {
        fence_after();
        if (MODE == 1) {
            const uint32_t idesc = make_idesc_f16(M, N, 0, 0);
#pragma unroll
            for (int s = 0; s < K / 16; s++)
                mma_f16_ss(tbase + d_col, make_smem_desc(smem_u32(a_s) + s * 256, 128, (K / 8) * 128), make_smem_desc(smem_u32(b_s) + s * 256, 128, (K / 8) * 128), idesc, s > 0);
        } else if (MODE == 2) {
            const uint32_t idesc = make_idesc_f16(M, N, 0, 1);
#pragma unroll
            for (int s = 0; s < K / 16; s++)
                mma_f16_ts(tbase + d_col, tbase + a_col + s * 8, make_smem_desc(smem_u32(b_s) + s * 256, 128, (K / 8) * 128), idesc, s > 0);
        } else {
            const uint32_t idesc = make_idesc_f16(M, N, 1, 1);
#pragma unroll
            for (int s = 0; s < K / 16; s++)
                mma_f16_ss(tbase + d_col, make_smem_desc(smem_u32(a_s) + s * 256, 128, (K / 8) * 128), make_smem_desc(smem_u32(b_s) + s * 256, 128, (K / 8) * 128), idesc, s > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after();
    uint32_t acc[16];
#pragma unroll
    for (int j = 0; j < N; j += 16) {
        tmem_ld16(tbase + lane_base + d_col + j, acc);
        wait_ld();
#pragma unroll
        for (int q = 0; q < 16; q++) D[(size_t)tid * N + j + q] = __uint_as_float(acc[q]);
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

template <int MODE, int M, int N, int K>
static int run() {
    std::vector<__half> A(M * K), B(N * K);
    srand(77 + MODE * 1000 + M + N * 3 + K * 5);
    auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.0f; };
    for (auto& v : A) v = __float2half(rnd());
    for (auto& v : B) v = __float2half(rnd() * 0.25f);
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe2<MODE, M, N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe2<MODE, M, N, K><<<1, 128, 65536>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE2 mode=%d M=%d N=%d K=%d CUDA ERROR %s\n", MODE, M, N, K, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            double acc = 0;
            for (int k = 0; k < K; k++) {
                float a = MODE == 3 ? __half2float(A[k * M + m]) : __half2float(A[m * K + k]);
                float b = MODE == 1 ? __half2float(B[n * K + k]) : __half2float(B[k * N + n]);
                acc += (double)a * b;
            }
            const int lane = M == 64 ? (m % 16) + 32 * (m / 16) : m;
            err = fmax(err, fabs(acc - D[lane * N + n]));
        }
    const bool ok = err < 2e-3;
    printf("PROBE2 mode=%d M=%d N=%d K=%d max_err=%.3e %s\n", MODE, M, N, K, err, ok ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run<1, 128, 64, 48>();
    bad += run<1, 128, 64, 80>();
    bad += run<1, 128, 64, 16>();
    bad += run<2, 128, 64, 64>();
    bad += run<2, 128, 64, 16>();
    bad += run<2, 128, 48, 64>();
    bad += run<2, 128, 80, 64>();
    bad += run<3, 64, 64, 128>();
    bad += run<3, 64, 48, 128>();
    bad += run<3, 64, 16, 128>();
    bad += run<3, 64, 80, 128>();
    printf(bad ? "TC05_PROBE2_FAIL\n" : "TC05_PROBE2_OK\n");
    return bad;
}
