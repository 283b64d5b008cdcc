// sm_100a kernels of the 128-neuron network (`nnWidth` = 128 on the reference's command line, src/AppConfig.cpp:169; tiny-cuda-nn
// instantiates FullyFusedMLP<__half, 128>, src/network.cu:117-118).  Same parameter order, encodings, loss and optimizer as the
// 64-neuron path (nrc_kernels.cuh); what changes is the shape of the tensor-core work:
//   * every hidden matrix is 128 x 128 fp16 = 32 KB, the whole network 176 KB for 48 inputs and six hidden layers: it still fits ONE
//     image in shared memory (no swizzle, canonical core-matrix layout, K-major for the forward pass, MN-major for the backward pass),
//     so the weights stay resident for the lifetime of a persistent CTA exactly as in the narrow network -- but one CTA per SM;
//   * tcgen05.mma tiles are M = 128 records x N = 128 neurons, K = 128 in eight K = 16 slices; a tile's fp32 accumulator takes 128
//     TMEM columns and its fp16 A operand (written by the epilogue with tcgen05.st, TS-mode chaining) 64: two tiles per CTA in flight;
//   * the activations of a training batch (6 x 128 x 128 fp16 = 192 KB per tile) do not fit next to the weights, so training runs as
//     forward / backward / weight-gradient kernels with the activations in HBM (the narrow network's fused step keeps them on chip);
//   * weight gradients: M = 128 rows of dW map one to one onto the 128 TMEM lanes (the narrow network needs the M = 64 lane mapping).
// Kernel arguments are the narrow path's (FwdArgs / BwdArgs / DwArgs) with [.][n][128] activation tensors.
#pragma once
#include "nrc_kernels.cuh"

namespace nrchpm {

constexpr int kWide = 128;                                   // n_neurons of this family
constexpr uint32_t kWideColsPerWg = 192;                     // 128 accumulator + 64 operand TMEM columns per warpgroup tile

template <int IN_W>
__host__ __device__ constexpr size_t wide_weight_bytes(int n_hidden) { return (size_t)IN_W * kWide * 2 + (size_t)(n_hidden - 1) * kWide * kWide * 2 + (size_t)kOutPad * kWide * 2; }
template <int IN_W>
__host__ __device__ constexpr size_t wide_fwd_smem_bytes(int n_hidden, int wgs) { return wide_weight_bytes<IN_W>(n_hidden) + (size_t)wgs * IN_W * 256; }
template <int IN_W>
__host__ __device__ constexpr size_t wide_bwd_smem_bytes(int n_hidden) { return wide_weight_bytes<IN_W>(n_hidden); }
constexpr size_t kWideDwSmemBytes = 2 * (2 * 32768);        // two stages of (A tile, B tile), 128 samples x 128 features fp16 each

// forward pass (inference, or the first kernel of a training step): one 128-record tile per warpgroup, one record per thread = one
// TMEM lane.  Replaces kernel_grid + kernel_one_blob + kernel_mlp_fused<128> (+ relative_l2_luminance_loss when TRAIN).
template <int IN_W, bool TRAIN>
__global__ void __launch_bounds__(256, 1) nrc_wide_forward_kernel(const __grid_constant__ FwdArgs a) {
    using namespace tc05;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float loss_red[2][4];
    const int tid = threadIdx.x, wg = tid >> 7, r = tid & 127, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x, nwg = nthreads >> 7;
    const int H = a.n_hidden;
    constexpr uint32_t kHid = kWide * kWide * 2, kSboW = (kWide / 8) * 128;      // bytes per hidden matrix; K-major stride between 8-row groups
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * kWide * 2;
    uint8_t* wo_s = wh_s + (size_t)(H - 1) * kHid;
    uint8_t* x_s = wo_s + kOutPad * kWide * 2 + wg * (IN_W * 256);

    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_mbar_init(); }
    copy_weights_kmajor(w0_s, a.params, kWide, IN_W, tid, nthreads);
    for (int l = 1; l < H; l++) copy_weights_kmajor(wh_s + (size_t)(l - 1) * kHid, a.params + IN_W * kWide + (size_t)(l - 1) * kWide * kWide, kWide, kWide, tid, nthreads);
    copy_weights_kmajor(wo_s, a.params + IN_W * kWide + (size_t)(H - 1) * kWide * kWide, kOutPad, kWide, tid, nthreads);
    fence_proxy_async_smem();
    fence_before();
    __syncthreads();
    fence_after();

    const uint32_t tD = tmem_base_s + wg * kWideColsPerWg, tA = tD + kWide;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc_h = make_idesc_f16(128, kWide), idesc_o = make_idesc_f16(128, kOutPad);
    const uint32_t x_addr = smem_u32(x_s), w0_addr = smem_u32(w0_s), wh_addr = smem_u32(wh_s), wo_addr = smem_u32(wo_s);
    const __half2* grid = reinterpret_cast<const __half2*>(a.params + a.n_mlp);
    uint64_t* bar = &mbar[wg];
    uint32_t phase = 0;

    uint32_t n = a.n;
    if (a.d_count) n = min(n, *a.d_count);
    const uint32_t n_tiles = (n + kTile - 1) / kTile;

    for (uint32_t tile = blockIdx.x * nwg + wg; tile < n_tiles; tile += gridDim.x * nwg) {
        const uint32_t row = tile * kTile + r;
        const bool valid = row < n;
        uint32_t rec = 0;
        float x0 = 0, x1 = 0, x2 = 0, th = 0, ph = 0;
        if (valid) {
            rec = a.indices ? a.indices[row] : row;
            const float* p = a.in + 5 * (size_t)rec;
            x0 = p[0]; x1 = p[1]; x2 = p[2]; th = p[3]; ph = p[4];
        }
        uint8_t* my_row = x_s + (r >> 3) * (IN_W * 16) + (r & 7) * 16;
        SmemRowPut put{my_row};
        encode_record<2>(a.enc, grid, x0, x1, x2, th, ph, put);
        fence_proxy_async_smem();
        fence_before();
        named_bar_sync(1 + wg, 128);
        if (r == 0) {
            fence_after();
#pragma unroll
            for (int s = 0; s < IN_W / 16; s++)
                mma_f16_ss(tD, make_smem_desc(x_addr + s * 256, 128, IN_W * 16), make_smem_desc(w0_addr + s * 256, 128, IN_W * 16), idesc_h, s > 0);
            mma_commit(bar);
        }
        if (TRAIN) {
            int4* dst = reinterpret_cast<int4*>(a.x16 + (size_t)row * IN_W);
#pragma unroll
            for (int c = 0; c < IN_W / 8; c++) dst[c] = *reinterpret_cast<const int4*>(my_row + c * 128);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();

        for (int l = 0; l < H; l++) {
            // epilogue of hidden layer l: ReLU -> fp16 -> A operand of the next layer (TMEM), 32 accumulator columns at a time
#pragma unroll
            for (int q = 0; q < kWide / 32; q++) {
                uint32_t acc[32], p[16];
                tmem_ld32(tD + lane_base + q * 32, acc);
                wait_ld();
                if (l == 0) {       // tcnn's ReLU is max(x, 0) in fp16: NaN -> 0; cvt.relu would keep it
                    const __half2 zero2 = __float2half2_rn(0.0f);
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        uint32_t v = pack_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                        __half2 m = __hmax2(*reinterpret_cast<__half2*>(&v), zero2);
                        p[j] = *reinterpret_cast<uint32_t*>(&m);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) p[j] = pack_relu_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                }
                tmem_st16(tA + lane_base + q * 16, p);
                if (TRAIN) {
                    int4* dst = reinterpret_cast<int4*>(a.acts + ((size_t)l * a.n + row) * kWide + q * 32);
#pragma unroll
                    for (int c = 0; c < 4; c++) dst[c] = make_int4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                }
            }
            wait_st();
            fence_before();
            named_bar_sync(1 + wg, 128);
            if (r == 0) {
                fence_after();
                if (l < H - 1) {
#pragma unroll
                    for (int s = 0; s < kWide / 16; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + l * kHid + s * 256, 128, kSboW), idesc_h, s > 0);
                } else {
#pragma unroll
                    for (int s = 0; s < kWide / 16; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wo_addr + s * 256, 128, kSboW), idesc_o, s > 0);
                }
                mma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
        }
        // output layer: fp16 like tcnn's network output, then float (common_device.h:990-999)
        uint32_t o[16];
        tmem_ld16(tD + lane_base, o);
        wait_ld();
        if (!TRAIN) {
            if (valid) {
                float* dst = a.out + 3 * (size_t)rec;
#pragma unroll
                for (int k = 0; k < 3; k++) dst[k] = __half2float(__float2half_rn(__uint_as_float(o[k])));
            }
        } else {
            uint32_t ph16[8];
#pragma unroll
            for (int j = 0; j < 8; j++) ph16[j] = pack_f16x2(__uint_as_float(o[2 * j]), __uint_as_float(o[2 * j + 1]));
            int4* od = reinterpret_cast<int4*>(a.out16 + (size_t)row * kOutPad);
            od[0] = make_int4(ph16[0], ph16[1], ph16[2], ph16[3]);
            od[1] = make_int4(ph16[4], ph16[5], ph16[6], ph16[7]);
            // RelativeL2Luminance (relative_l2_luminance.h:40-88); n_total = batch * 3 (padded dims contribute nothing)
            float pr[3];
#pragma unroll
            for (int k = 0; k < 3; k++) pr[k] = __half2float(__float2half_rn(__uint_as_float(o[k])));
            const float n_total = (float)(a.n * 3u);
            const float lum = 0.299f * pr[0] + 0.587f * pr[1] + 0.114f * pr[2];
            const float denom = lum * lum + 0.01f;
            float loss = 0, gk[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float diff = pr[k] - a.target[3 * (size_t)row + k];
                loss += diff * diff / denom / n_total;
                gk[k] = a.loss_scale * (2 * diff / denom) / n_total;
            }
            int4* gd = reinterpret_cast<int4*>(a.dout16 + (size_t)row * kOutPad);
            gd[0] = make_int4(pack_f16x2(gk[0], gk[1]), pack_f16x2(gk[2], 0.0f), 0, 0);
            gd[1] = make_int4(0, 0, 0, 0);
            // deterministic per-tile loss sum: warp shuffle tree, then 4 partials added in order
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, s);
            if (lane == 0) loss_red[wg][warp & 3] = loss;
            named_bar_sync(1 + wg, 128);
            if (r == 0) a.loss_partials[tile] = ((loss_red[wg][0] + loss_red[wg][1]) + loss_red[wg][2]) + loss_red[wg][3];
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// backward pass through the network and into the hash grid.  Replaces kernel_mlp_fused_backward<128> + fc_multiply (dL/dinput) +
// kernel_grid_backward; weights are copied UNtransposed and read as MN-major B operands (B[n = input][k = output]).
template <int IN_W>
__global__ void __launch_bounds__(256, 1) nrc_wide_backward_kernel(const __grid_constant__ BwdArgs a) {
    using namespace tc05;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, wg = tid >> 7, r = tid & 127, warp = tid >> 5;
    const int nthreads = blockDim.x, nwg = nthreads >> 7;
    const int H = a.n_hidden;
    constexpr uint32_t kHid = kWide * kWide * 2, kSboW = (kWide / 8) * 128;      // MN-major: stride between 8-input groups
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * kWide * 2;
    uint8_t* wo_s = wh_s + (size_t)(H - 1) * kHid;

    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_mbar_init(); }
    if (blockIdx.x == 0 && tid == 0 && a.loss_out) {     // Trainer::loss (trainer.h:205-207): fixed-order sum of the tile partials
        float s = 0;
        for (uint32_t i = 0; i < a.n_loss_partials; i++) s += a.loss_partials[i];
        *a.loss_out = s;
    }
    copy_weights_mnmajor(w0_s, a.params, kWide, IN_W, tid, nthreads);
    for (int l = 1; l < H; l++) copy_weights_mnmajor(wh_s + (size_t)(l - 1) * kHid, a.params + IN_W * kWide + (size_t)(l - 1) * kWide * kWide, kWide, kWide, tid, nthreads);
    copy_weights_mnmajor(wo_s, a.params + IN_W * kWide + (size_t)(H - 1) * kWide * kWide, kOutPad, kWide, tid, nthreads);
    fence_proxy_async_smem();
    fence_before();
    __syncthreads();
    fence_after();

    const uint32_t tD = tmem_base_s + wg * kWideColsPerWg, tA = tD + kWide;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc_h = make_idesc_f16(128, kWide, 0, 1), idesc_x = make_idesc_f16(128, IN_W, 0, 1);
    const uint32_t w0_addr = smem_u32(w0_s), wh_addr = smem_u32(wh_s), wo_addr = smem_u32(wo_s);
    uint64_t* bar = &mbar[wg];
    uint32_t phase = 0;
    const uint32_t n_tiles = a.n / kTile;

    for (uint32_t tile = blockIdx.x * nwg + wg; tile < n_tiles; tile += gridDim.x * nwg) {
        const uint32_t row = tile * kTile + r;
        {   // dL/doutput row -> A operand (K = 16)
            const int4* src = reinterpret_cast<const int4*>(a.dout16 + (size_t)row * kOutPad);
            const int4 v0 = src[0], v1 = src[1];
            uint32_t p[8] = {(uint32_t)v0.x, (uint32_t)v0.y, (uint32_t)v0.z, (uint32_t)v0.w, (uint32_t)v1.x, (uint32_t)v1.y, (uint32_t)v1.z, (uint32_t)v1.w};
            tmem_st8(tA + lane_base, p);
        }
        wait_st();
        fence_before();
        named_bar_sync(1 + wg, 128);
        if (r == 0) {
            fence_after();
            // dA_{H-1} = dOut (K = 16 outputs) x W_out[out][in] read as B[n = in][k = out]: 8-input groups are (16 / 8) * 128 bytes apart
            mma_f16_ts(tD, tA, make_smem_desc(wo_addr, 128, (kOutPad / 8) * 128), idesc_h, 0);
            mma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();

        for (int l = H - 1; l >= 0; l--) {
            const int4* ap = reinterpret_cast<const int4*>(a.acts + ((size_t)l * a.n + row) * kWide);
            int4* dp = reinterpret_cast<int4*>(a.dacts + ((size_t)l * a.n + row) * kWide);
#pragma unroll
            for (int q = 0; q < kWide / 32; q++) {
                uint32_t acc[32], p[16];
                int4 av[4];
#pragma unroll
                for (int c = 0; c < 4; c++) av[c] = ap[q * 4 + c];
                tmem_ld32(tD + lane_base + q * 32, acc);
                wait_ld();
                const uint32_t* aw = reinterpret_cast<const uint32_t*>(av);
#pragma unroll
                for (int j = 0; j < 16; j++) {       // ReLU mask from the stored activations
                    const float lo = (aw[j] & 0x0000ffffu) ? __uint_as_float(acc[2 * j]) : 0.0f;
                    const float hi = (aw[j] & 0xffff0000u) ? __uint_as_float(acc[2 * j + 1]) : 0.0f;
                    p[j] = pack_f16x2(lo, hi);
                }
                tmem_st16(tA + lane_base + q * 16, p);
#pragma unroll
                for (int c = 0; c < 4; c++) dp[q * 4 + c] = make_int4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
            }
            if (l == 0 && !a.need_dx) break;
            wait_st();
            fence_before();
            named_bar_sync(1 + wg, 128);
            if (r == 0) {
                fence_after();
                if (l > 0) {
#pragma unroll
                    for (int s = 0; s < kWide / 16; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + (l - 1) * kHid + s * 256, 128, kSboW), idesc_h, s > 0);
                } else {
#pragma unroll
                    for (int s = 0; s < kWide / 16; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(w0_addr + s * 256, 128, kSboW), idesc_x, s > 0);
                }
                mma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
        }
        if (a.need_dx) {
            // dL/d(network input): fp16 like tcnn's fc_multiply output (fully_fused_mlp.cu:832-835)
            uint32_t acc[32];
            tmem_ld32(tD + lane_base, acc);
            wait_ld();
            if (a.dx16) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(a.dx16 + (size_t)row * IN_W);
#pragma unroll
                for (int j = 0; j < (IN_W < 32 ? IN_W / 2 : 16); j++) dst[j] = pack_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
#pragma unroll
                for (int c = 32; c < IN_W; c += 16) {
                    uint32_t t[16];
                    tmem_ld16(tD + lane_base + c, t);
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; j++) dst[c / 2 + j] = pack_f16x2(__uint_as_float(t[2 * j]), __uint_as_float(t[2 * j + 1]));
                }
            }
            if (a.grid_grad && a.enc.pos_enc == POS_HASHGRID) {
                // kernel_grid_backward (grid.h:215-320): (half)weight * grad, fp16x2 reductions
                const float* rp = a.in + 5 * (size_t)row;
                const float x0 = rp[0], x1 = rp[1], x2 = rp[2];
                __half2* gg = reinterpret_cast<__half2*>(a.grid_grad);
#pragma unroll
                for (int l = 0; l < kMaxLevels; l++) {
                    if (l >= a.enc.n_levels) break;
                    uint32_t gp = pack_f16x2(__uint_as_float(acc[2 * l]), __uint_as_float(acc[2 * l + 1]));
                    const __half2 g = *reinterpret_cast<__half2*>(&gp);
                    GridLevel c;
                    grid_level_cell(a.enc, l, x0, x1, x2, c);
                    __half2* base = gg + a.enc.level_offset[l];
#pragma unroll
                    for (int k = 0; k < 8; k++) red_add_f16x2(base + c.idx[k], __hmul2(__half2half2(__float2half_rn(c.w[k])), g));
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// weight gradients dW_m = sum over the batch of dY_m^T A_{m-1}: the batch is the contraction dimension, both operands MN-major
// ([sample][feature] rows copied as they lie).  blockIdx.x = batch chunk, blockIdx.y = weight matrix; fp32 partial per chunk, added
// in a fixed order by the optimizer (deterministic, no atomics).  Replaces the split-K CUTLASS GEMMs of fully_fused_mlp.cu:783-836.
template <int IN_W>
__global__ void __launch_bounds__(128, 1) nrc_wide_dw_kernel(const __grid_constant__ DwArgs a) {
    using namespace tc05;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int H = a.n_hidden, m = blockIdx.y;
    // operands of matrix m: A-op = the 128 rows of D, B-op = its N columns
    const __half* aop; const __half* bop; int bw;
    if (m == 0) { aop = a.dacts; bop = a.x16; bw = IN_W; }
    else if (m < H) { aop = a.dacts + (size_t)m * a.n * kWide; bop = a.acts + (size_t)(m - 1) * a.n * kWide; bw = kWide; }
    else { aop = a.acts + (size_t)(H - 1) * a.n * kWide; bop = a.dout16; bw = kOutPad; }      // output layer: computed transposed, D[in][out]

    if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_mbar_init(); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tD = tmem_base_s;
    const uint32_t idesc = make_idesc_f16(128, (uint32_t)bw, 1, 1);
    const uint32_t b0 = blockIdx.x * a.kc, b1 = min(a.n, b0 + a.kc);
    uint32_t phase[2] = {0, 0};
    int it = 0;
    for (uint32_t b = b0; b < b1; b += kTile, it++) {
        const int buf = it & 1;
        uint8_t* as = smem + buf * 65536;
        uint8_t* bs = as + 32768;
        if (it >= 2) { mbar_wait(&mbar[buf], phase[buf]); phase[buf] ^= 1; }     // the MMAs that read this stage are done
        // thread = sample: its 16-byte chunk of feature group c lands at c * 2048 + (sample % 8) * 16 + (sample / 8) * 128
        const uint32_t off = (tid & 7) * 16 + (tid >> 3) * 128;
        const int4* ag = reinterpret_cast<const int4*>(aop + (size_t)(b + tid) * kWide);
#pragma unroll
        for (int c = 0; c < kWide / 8; c++) *reinterpret_cast<int4*>(as + c * 2048 + off) = ag[c];
        const int4* bg = reinterpret_cast<const int4*>(bop + (size_t)(b + tid) * bw);
        for (int c = 0; c < bw / 8; c++) *reinterpret_cast<int4*>(bs + c * 2048 + off) = bg[c];
        fence_proxy_async_smem();
        fence_before();
        __syncthreads();
        if (tid == 0) {
            fence_after();
            const uint32_t aa = smem_u32(as), ba = smem_u32(bs);
#pragma unroll
            for (int s = 0; s < 8; s++) mma_f16_ss(tD, make_smem_desc(aa + s * 256, 128, 2048), make_smem_desc(ba + s * 256, 128, 2048), idesc, (it > 0 || s > 0) ? 1u : 0u);
            mma_commit(&mbar[buf]);
        }
    }
    {   // drain: the last commit covers every earlier MMA
        const int last = (it - 1) & 1;
        mbar_wait(&mbar[last], phase[last]);
        fence_after();
    }
    // M = 128: row `tid` of D lives in TMEM lane `tid`
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* out = a.partials + (size_t)blockIdx.x * a.n_mlp;
    const size_t moff = m == 0 ? 0 : (size_t)IN_W * kWide + (size_t)(m - 1) * kWide * kWide;
    for (int c = 0; c < bw; c += 16) {
        uint32_t v[16];
        tmem_ld16(tD + lane_base + c, v);
        wait_ld();
        if (m < H) {
            float4* dst = reinterpret_cast<float4*>(out + moff + (size_t)tid * bw + c);
#pragma unroll
            for (int q = 0; q < 4; q++) dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        } else {      // D[in][out] -> W_out[out][in]
#pragma unroll
            for (int q = 0; q < 16; q++) out[moff + (size_t)(c + q) * kWide + tid] = __uint_as_float(v[q]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 128);
}

// warp-specialised inference (the 128-neuron counterpart of nrc_infer_ws_kernel): NP producer warpgroups encode records into a ring
// of NS K-major X tiles in shared memory, NC consumer warpgroups run the seven-layer MLP (tcgen05, activations in TMEM).  With one
// tile per warpgroup (nrc_wide_forward_kernel) a 1080p frame of hash-grid records takes 1.58 ms: eight warps per SM cannot keep the
// gathers of 16 levels in flight.  Two consumers sustain 0.55 ms per frame (the M = 128 x N = 128 layers are 4x the narrow
// network's work per tile), four producers keep the L1 -> L2 request path busy underneath.
#ifndef NRC_WIDE_WS_NP
#define NRC_WIDE_WS_NP 4
#endif
#ifndef NRC_WIDE_WS_NC
#define NRC_WIDE_WS_NC 2
#endif
template <int IN_W>
__host__ __device__ constexpr int wide_ws_slots(int n_hidden) {      // ring slots that fit next to the weight image (227 KB per CTA), at most 4
    return (int)((227 * 1024 - 2048 - wide_weight_bytes<IN_W>(n_hidden)) / ((size_t)IN_W * 256)) > 4 ? 4 : (int)((227 * 1024 - 2048 - wide_weight_bytes<IN_W>(n_hidden)) / ((size_t)IN_W * 256));
}
template <int IN_W>
__host__ __device__ constexpr size_t wide_ws_smem_bytes(int n_hidden, int slots) { return wide_weight_bytes<IN_W>(n_hidden) + (size_t)slots * IN_W * 256; }

template <int IN_W, int NP, int NC>
__global__ void __maxnreg__(80) nrc_wide_infer_ws_kernel(const __grid_constant__ FwdArgs a) {
    using namespace tc05;
    constexpr int kMaxSlots = 4;
    const uint32_t NS = a.ring_slots;
    constexpr uint32_t kAlloc = NC * kWideColsPerWg <= 256 ? 256u : 512u;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[kMaxSlots], empty[kMaxSlots], mbar[NC];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, r = tid & 127;
    const int H = a.n_hidden;
    constexpr uint32_t kHid = kWide * kWide * 2, kSboW = (kWide / 8) * 128;
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * kWide * 2;
    uint8_t* wo_s = wh_s + (size_t)(H - 1) * kHid;
    uint8_t* ring = wo_s + kOutPad * kWide * 2;

    timeline_begin(a.tl, 0);
    if (warp == 0) { tmem_alloc(&tmem_base_s, kAlloc); tmem_relinquish(); }
    if (tid == 0) {
        for (uint32_t s = 0; s < NS; s++) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
        for (int c = 0; c < NC; c++) mbar_init(&mbar[c], 1);
        fence_mbar_init();
    }
    copy_weights_kmajor(w0_s, a.params, kWide, IN_W, tid, (NP + NC) * 128);
    for (int l = 1; l < H; l++) copy_weights_kmajor(wh_s + (size_t)(l - 1) * kHid, a.params + IN_W * kWide + (size_t)(l - 1) * kWide * kWide, kWide, kWide, tid, (NP + NC) * 128);
    copy_weights_kmajor(wo_s, a.params + IN_W * kWide + (size_t)(H - 1) * kWide * kWide, kOutPad, kWide, tid, (NP + NC) * 128);
    fence_proxy_async_smem();
    fence_before();
    __syncthreads();
    fence_after();

    uint32_t n = a.n;
    if (a.d_count) n = min(n, *a.d_count);
    const uint32_t n_tiles = (n + kTile - 1) / kTile;
    const uint32_t n_my = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;      // tiles blockIdx.x + k * gridDim.x

    if (wg < NP) {
        // ------------------------------------------------------------ producers: records -> encoded X tiles
        const __half2* grid = reinterpret_cast<const __half2*>(a.params + a.n_mlp);
        const bool by_kind = a.enc.pos_enc == POS_HASHGRID && a.enc.all_pow2;
        for (uint32_t k = wg; k < n_my; k += NP) {
            const uint32_t slot = k % NS, use = k / NS;
            const uint32_t row = (blockIdx.x + k * gridDim.x) * kTile + r;
            float x0 = 0, x1 = 0, x2 = 0, th = 0, ph = 0;
            if (row < n) {
                const uint32_t rec = a.indices ? a.indices[row] : row;
                const float* p = a.in + 5 * (size_t)rec;
                x0 = __ldcs(p); x1 = __ldcs(p + 1); x2 = __ldcs(p + 2); th = __ldcs(p + 3); ph = __ldcs(p + 4);      // read once: keep them out of L1
            }
            if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);          // the tensor core has read the tile that lived here
            SmemRowPut put{ring + slot * (IN_W * 256) + (r >> 3) * (IN_W * 16) + (r & 7) * 16};
            if (by_kind) { hashgrid_by_kind(a.enc, grid, x0, x1, x2, put); encode_direction_pad(a.enc, th, ph, put); }
            else encode_record<1>(a.enc, grid, x0, x1, x2, th, ph, put);
            fence_proxy_async_smem();
            mbar_arrive(&full[slot]);
        }
    } else {
        // ------------------------------------------------------------ consumers: the MLP
        const int c = wg - NP;
        const uint32_t tD = tmem_base_s + c * kWideColsPerWg, tA = tD + kWide;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t idesc_h = make_idesc_f16(128, kWide), idesc_o = make_idesc_f16(128, kOutPad);
        const uint32_t ring_addr = smem_u32(ring), w0_addr = smem_u32(w0_s), wh_addr = smem_u32(wh_s), wo_addr = smem_u32(wo_s);
        uint64_t* bar = &mbar[c];
        uint32_t phase = 0;
        for (uint32_t k = c; k < n_my; k += NC) {
            const uint32_t slot = k % NS, use = k / NS;
            const uint32_t row = (blockIdx.x + k * gridDim.x) * kTile + r;
            const bool valid = row < n;
            uint32_t rec = 0;
            if (valid) rec = a.indices ? a.indices[row] : row;
            mbar_wait(&full[slot], use & 1);
            if (r == 0) {
                fence_after();
                const uint32_t x_addr = ring_addr + slot * (IN_W * 256);
#pragma unroll
                for (int s = 0; s < IN_W / 16; s++)
                    mma_f16_ss(tD, make_smem_desc(x_addr + s * 256, 128, IN_W * 16), make_smem_desc(w0_addr + s * 256, 128, IN_W * 16), idesc_h, s > 0);
                mma_commit(&empty[slot]);       // the slot returns to the producers as soon as the tensor core has read it
                mma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
            for (int l = 0; l < H; l++) {
#pragma unroll
                for (int q = 0; q < kWide / 32; q++) {
                    uint32_t acc[32], p[16];
                    tmem_ld32(tD + lane_base + q * 32, acc);
                    wait_ld();
                    if (l == 0) {      // tcnn's ReLU is max(x, 0) in fp16: NaN (Q5) -> 0; cvt.relu would keep it
                        const __half2 zero2 = __float2half2_rn(0.0f);
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            uint32_t v = pack_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                            __half2 m = __hmax2(*reinterpret_cast<__half2*>(&v), zero2);
                            p[j] = *reinterpret_cast<uint32_t*>(&m);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j++) p[j] = pack_relu_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                    }
                    tmem_st16(tA + lane_base + q * 16, p);
                }
                wait_st();
                fence_before();
                named_bar_sync(1 + c, 128);
                if (r == 0) {
                    fence_after();
                    if (l < H - 1) {
#pragma unroll
                        for (int s = 0; s < kWide / 16; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + l * kHid + s * 256, 128, kSboW), idesc_h, s > 0);
                    } else {
#pragma unroll
                        for (int s = 0; s < kWide / 16; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wo_addr + s * 256, 128, kSboW), idesc_o, s > 0);
                    }
                    mma_commit(bar);
                }
                mbar_wait(bar, phase); phase ^= 1;
                fence_after();
            }
            // output layer: fp16 like tcnn's network output, then float (common_device.h:990-999)
            uint32_t o[4];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "r"(tD + lane_base) : "memory");
            wait_ld();
            if (valid) {
                float* dst = a.out + 3 * (size_t)rec;
#pragma unroll
                for (int q = 0; q < 3; q++) __stcs(dst + q, __half2float(__float2half_rn(__uint_as_float(o[q]))));
            }
            fence_before();
            named_bar_sync(1 + c, 128);      // every thread's read of tD is complete before thread 0 issues the next tile's layer-0 MMA
        }
    }
    fence_before();
    __syncthreads();
    timeline_end(a.tl, 0);
    if (warp == 0) tmem_dealloc(tmem_base_s, kAlloc);
}

}  // namespace nrchpm
