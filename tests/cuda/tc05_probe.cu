// Stand-alone probe for the tcgen05 building blocks the NRC MLP relies on (run on the B200 via
// tests/test_gpu_tc05_probe.py): A operand staged in TMEM with tcgen05.st (one row per thread, fp16
// pairs packed along columns), B operand = row-major [N][K] weights in shared memory in the
// no-swizzle K-major canonical layout, fp32 accumulators read back with tcgen05.ld.
// Checks  D = A * W^T  (K = 64 and K = 48, N = 64 and N = 16) and a chained second layer.
#include "../../nrc_hpm_renderer_b200/csrc/tc05.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

using namespace tc05;

template <int K, int N>
__global__ void __launch_bounds__(128, 1) probe_kernel(const __half* __restrict__ A, const __half* __restrict__ W, const __half* __restrict__ W2,
                                                       float* __restrict__ D, float* __restrict__ D2) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* w_s = smem;                       // [N][K] K-major
    uint8_t* w2_s = smem + 16384;              // [16][N] K-major (second layer: K2 = N, N2 = 16)
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * K; i += 128) {
        int n = i / K, k = i % K;
        *reinterpret_cast<__half*>(w_s + kmajor_offset(n, k, K)) = W[i];
    }
    for (int i = tid; i < 16 * N; i += 128) {
        int n = i / N, k = i % N;
        *reinterpret_cast<__half*>(w2_s + kmajor_offset(n, k, N)) = W2[i];
    }
    fence_proxy_async_smem();
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t d_col = 0, a_col = 64;      // D: cols 0..63, A: cols 64..(64+K/2)

    // stage my row of A into TMEM
    uint32_t a[32];
#pragma unroll
    for (int j = 0; j < K / 2; j++) a[j] = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K)[j];
#pragma unroll
    for (int j = 0; j < K / 2; j += 8) tmem_st8(tbase + lane_base + a_col + j, a + j);
    wait_st();
    fence_before();
    __syncthreads();
    if (tid == 0) {
        fence_after();
        const uint32_t idesc = make_idesc_f16(128, N);
#pragma unroll
        for (int s = 0; s < K / 16; s++) {
            uint64_t bd = make_smem_desc(smem_u32(w_s) + s * 256, 128, (K / 8) * 128);
            mma_f16_ts(tbase + d_col, tbase + a_col + s * 8, bd, idesc, s > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after();
    uint32_t acc[64];
#pragma unroll
    for (int j = 0; j < N; j += 16) tmem_ld16(tbase + lane_base + d_col + j, acc + j);
    wait_ld();
#pragma unroll
    for (int j = 0; j < N; j++) D[(size_t)tid * N + j] = __uint_as_float(acc[j]);

    // second layer: relu -> fp16 -> back into the A region -> N2 = 16
    uint32_t p[32];
#pragma unroll
    for (int j = 0; j < N / 2; j++) p[j] = pack_relu_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
#pragma unroll
    for (int j = 0; j < N / 2; j += 8) tmem_st8(tbase + lane_base + a_col + j, p + j);
    wait_st();
    fence_before();
    __syncthreads();
    if (tid == 0) {
        fence_after();
        const uint32_t idesc = make_idesc_f16(128, 16);
#pragma unroll
        for (int s = 0; s < N / 16; s++) {
            uint64_t bd = make_smem_desc(smem_u32(w2_s) + s * 256, 128, (N / 8) * 128);
            mma_f16_ts(tbase + d_col, tbase + a_col + s * 8, bd, idesc, s > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 1);
    fence_after();
    uint32_t o[16];
    tmem_ld16(tbase + lane_base + d_col, o);
    wait_ld();
#pragma unroll
    for (int j = 0; j < 16; j++) D2[(size_t)tid * 16 + j] = __uint_as_float(o[j]);
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

template <int K, int N>
static int run() {
    std::vector<__half> A(128 * K), W(N * K), W2(16 * N);
    srand(1234 + K * 7 + N);
    auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.0f; };
    for (auto& v : A) v = __float2half(rnd());
    for (auto& v : W) v = __float2half(rnd() * 0.25f);
    for (auto& v : W2) v = __float2half(rnd() * 0.25f);
    __half *dA, *dW, *dW2; float *dD, *dD2;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dW, W.size() * 2); cudaMalloc(&dW2, W2.size() * 2);
    cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dD2, 128 * 16 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dW2, W2.data(), W2.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    probe_kernel<K, N><<<1, 128, 32768>>>(dA, dW, dW2, dD, dD2);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("PROBE K=%d N=%d CUDA ERROR %s\n", K, N, cudaGetErrorString(e)); return 1; }
    std::vector<float> D(128 * N), D2(128 * 16);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int m = 0; m < 128; m++) {
        std::vector<float> h1(N);
        for (int n = 0; n < N; n++) {
            double acc = 0;
            for (int k = 0; k < K; k++) acc += (double)__half2float(A[m * K + k]) * (double)__half2float(W[n * K + k]);
            e1 = fmax(e1, fabs(acc - D[m * N + n]));
            h1[n] = __half2float(__float2half(fmaxf(D[m * N + n], 0.0f)));
        }
        for (int n = 0; n < 16; n++) {
            double acc = 0;
            for (int k = 0; k < N; k++) acc += (double)h1[k] * (double)__half2float(W2[n * N + k]);
            e2 = fmax(e2, fabs(acc - D2[m * 16 + n]));
        }
    }
    const bool ok = e1 < 1e-3 && e2 < 1e-3;
    printf("PROBE K=%d N=%d max_err_layer1=%.3e max_err_layer2=%.3e %s\n", K, N, e1, e2, ok ? "OK" : "FAIL");
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run<64, 64>();
    bad += run<48, 64>();
    bad += run<16, 64>();
    bad += run<64, 48>();
    printf(bad ? "TC05_PROBE_FAIL\n" : "TC05_PROBE_OK\n");
    return bad;
}
