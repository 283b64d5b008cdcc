// sm_100a kernels of the Neural Radiance Cache: input encodings, fully fused MLP forward (inference and training),
// loss, fused backward (+ hash-grid gradient scatter), weight gradients, Adam + EMA.
//
// What they replace in the reference (all in the tiny-cuda-nn submodule, paths relative to tiny-cuda-nn/):
//   include/tiny-cuda-nn/encodings/grid.h:49-212 (kernel_grid), :215-320 (kernel_grid_backward)
//   include/tiny-cuda-nn/encodings/oneblob.h:85-127, triangle_wave.h:46-82, frequency.h:46-80, identity.h:46-66
//   src/fully_fused_mlp.cu:499-557 (kernel_mlp_fused), :150-259 (kernel_mlp_fused_backward), :783-836 (weight / input
//   gradient GEMMs through CUTLASS), include/tiny-cuda-nn/common_device.h:990-999 (trim_and_cast)
//   include/tiny-cuda-nn/losses/relative_l2_luminance.h:40-88, optimizers/adam.h:48-121, optimizers/ema.h:63-76
//
// Design (B200-first, not a translation): one warpgroup (128 threads) owns a tile of 128 records, one record per
// thread == one TMEM lane.  The activations never leave the SM: layer 0 reads the encoded inputs from shared memory
// (tcgen05.mma SS), every later layer takes its A operand from TMEM (tcgen05.mma TS) where the previous epilogue
// (tcgen05.ld -> ReLU -> fp16 -> tcgen05.st) left it.  All weight matrices stay resident in shared memory in the
// canonical no-swizzle UMMA layout (K-major for forward, MN-major for backward, so the row-major fp16 parameter
// vector is copied in 16-byte chunks without any transpose).  fp32 accumulation in TMEM (tcnn accumulates in fp16).
#pragma once
#include <cuda_fp16.h>
#include <cstdint>
#include "tc05.cuh"
#include "nrc_types.h"

// width of the hash-grid gather loads: 64 (aligned pair of entries) or 128 (aligned group of four entries)
#ifndef NRC_GATHER_BITS
#define NRC_GATHER_BITS 64
#endif

// cache operator of the hash-grid gather loads: 0 = default (L1-allocating), 1 = ld.global.cg (L2 only)
#ifndef NRC_GATHER_CG
#define NRC_GATHER_CG 0
#endif

// 1: software-pipelined hash-grid gathers (level l+1 in flight while level l is interpolated); 0: rounds of NRC_INFER_UNROLL levels.
// Measured on B200, 2 073 600 random records (profiles/r01_summary.md): rounds of ONE level 0.636 ms, rounds of two 0.720 ms,
// pipelined 0.698 ms, 128-bit groups 0.664 ms -- the more gathers are in flight, the more 128-byte lines the small L1 (84 KB next
// to 144 KB of shared memory) has to hold and the fewer second-corner loads still find the sector of their twin (L1 hit rate 30 %
// -> 15 %), which turns into extra L1->L2 requests, the unit this kernel saturates first.  Hence the default: one level per round.
#ifndef NRC_GATHER_PIPE
#define NRC_GATHER_PIPE 0
#endif
#ifndef NRC_TRAIN_UNROLL
#define NRC_TRAIN_UNROLL 2           // levels per round in the (latency-bound, low-occupancy) training forward kernel
#endif
#ifndef NRC_TRAIN_PREFETCH
#define NRC_TRAIN_PREFETCH 0         // fused training kernel: prefetch the Adam records of the touched hash-grid entries into L2
#endif

namespace nrchpm {

template <class T>
__device__ __forceinline__ T gather_ld(const void* p) {
#if NRC_GATHER_CG
    return __ldcg(reinterpret_cast<const T*>(p));
#else
    return *reinterpret_cast<const T*>(p);
#endif
}

// ---------------------------------------------------------------------------------------------- encodings
__device__ __forceinline__ float quartic_cdf(float x, float inv_radius) {      // common_device.h:905-920
    const float u = x * inv_radius;
    const float u2 = u * u;
    const float u4 = u2 * u2;
    return fmaxf(0.0f, fminf(1.0f, ((float)15 / 16) * u * (1 - ((float)2 / 3) * u2 + ((float)1 / 5) * u4) + 0.5f));
}

struct GridLevel {
    uint32_t idx[8];
    float w[8];
};

// grid.h:49-212 / common_device.h:632-718, 842-868: cell corners (entry indices relative to the level) and weights
template <bool POW2 = false>
__device__ __forceinline__ void grid_level_cell(const EncParams& e, int l, float x0, float x1, float x2, GridLevel& c) {
    const float scale = e.level_scale[l];
    float p[3] = {fmaf(scale, x0, 0.5f), fmaf(scale, x1, 0.5f), fmaf(scale, x2, 0.5f)};
    uint32_t g[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float t = floorf(p[d]);
        g[d] = (uint32_t)(int)t;
        p[d] -= t;
    }
    const uint32_t hs = e.level_hsize[l], s0 = e.level_s0[l], s1 = e.level_s1[l], s2 = e.level_s2[l];
    const bool hashed = e.level_hash[l] != 0;
    // `% hashmap_size` is a mask for power-of-two tables (every level of the presets: POW2 instantiation, no division code)
    const bool pow2 = POW2 || (hs & (hs - 1)) == 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float w = 1;
        uint32_t q[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if ((k & (1 << d)) == 0) { w *= 1 - p[d]; q[d] = g[d]; } else { w *= p[d]; q[d] = g[d] + 1; }
        }
        const uint32_t ih = (q[0] * 1u) ^ (q[1] * 2654435761u) ^ (q[2] * 805459861u), il = q[0] * s0 + q[1] * s1 + q[2] * s2;
        const uint32_t index = hashed ? ih : il;
        if (POW2) c.idx[k] = index & (hs - 1);
        else c.idx[k] = pow2 ? (index & (hs - 1)) : (index % hs);
        c.w[k] = w;
    }
}

// one hash-grid level whose table has been staged in shared memory (inference kernel, coarse levels): the same corners, weights
// and fp16 fma order as the global-memory path below; a divergent LDS costs its bank-conflict degree (~3-4 cycles per warp)
// instead of one L1 wavefront per lane
__device__ __forceinline__ __half2 encode_level_smem(const EncParams& e, int l, uint32_t table_smem_addr, float x0, float x1, float x2) {
    GridLevel c;
    grid_level_cell(e, l, x0, x1, x2, c);
    __half2 r = __float2half2_rn(0.0f);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(table_smem_addr + c.idx[k] * 4u));
        r = __hfma2(__half2half2(__float2half_rn(c.w[k])), *reinterpret_cast<const __half2*>(&v), r);
    }
    return r;
}

// Hash-grid levels [l_begin, l_end), UNROLL levels per round: every gather of a round is issued before the first one is
// consumed (the kernel is bound by the latency of these L2 round trips), so the body is branch-free -- one basic block of
// predicated loads followed by one of fp16 fmas.  Gather diet (bit-exact):
//  * a corner whose trilinear weight is exactly 0 contributes fma(0, v, r) == r, so its load is predicated off.  With the
//    reference's position normalisation (coordinates ~30) the fractional part is 0 in every dimension at the
//    finest level and often at the next ones: ~13 % of all gathers disappear;
//  * the two corners of an x-edge sit in ONE aligned 8-byte word (16-byte group with NRC_GATHER_BITS == 128) whenever their
//    indices differ only in the low bit(s) (dense levels and the coherent-prime hash, whose x prime is 1) -> one wide load
//    instead of two 32-bit loads, i.e. one L1 wavefront instead of two for half (three quarters) of all edges;
//  * tcnn's stride arithmetic wraps in uint32 for resolutions >= 2^16 (common_device.h:842-868): the z stride (and at the
//    finest level the y stride too) is 0 modulo the table size, so corners k and k+4 (k and k+2) are the SAME entry.  The
//    aliasing corners are never loaded; they take the value of their twin through a select.
template <int UNROLL, bool POW2, class Put>
__device__ __forceinline__ void hashgrid_levels(const EncParams& e, const __half2* __restrict__ grid, float x0, float x1, float x2,
                                                int l_begin, int l_end, Put& put, const GridAdamState* pf_state = nullptr) {
#if NRC_GATHER_BITS == 128
    typedef uint4 wide_t;
    constexpr uint32_t kGroupMask = 3u;
#else
    typedef uint2 wide_t;
    constexpr uint32_t kGroupMask = 1u;
#endif
    for (int l0 = l_begin; l0 < l_end; l0 += UNROLL) {
        GridLevel c[UNROLL];
        wide_t q[UNROLL][4];
        uint32_t a1[UNROLL][4];
        bool dupz[UNROLL], dupy[UNROLL];
        // ---- addresses and predicated loads of the whole round
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const bool on = l0 + u < l_end;
            const int l = on ? l0 + u : l_end - 1;
            grid_level_cell<POW2>(e, l, x0, x1, x2, c[u]);
            const __half2* base = grid + e.level_offset[l];
            if (pf_state) {      // training: the optimizer will update exactly these entries -- pull their Adam records into L2 (fire-and-forget)
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (on) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_state + e.level_offset[l] + c[u].idx[k]));
            }
            const uint32_t hs_m = e.level_hsize[l] - 1u;
            const bool lin = e.level_hash[l] == 0 && (e.level_hsize[l] & hs_m) == 0;
            dupz[u] = lin && (e.level_s2[l] & hs_m) == 0;
            dupy[u] = dupz[u] && (e.level_s1[l] & hs_m) == 0;
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const int j = k >> 1;
                const uint32_t i0 = c[u].idx[k], i1 = c[u].idx[k + 1];
                bool n0 = c[u].w[k] != 0.0f, n1 = c[u].w[k + 1] != 0.0f;
                if (k < 4) { n0 |= dupz[u] & (c[u].w[k + 4] != 0.0f); n1 |= dupz[u] & (c[u].w[k + 5] != 0.0f); }
                if (k == 0) { n0 |= dupy[u] & ((c[u].w[2] != 0.0f) | (c[u].w[6] != 0.0f)); n1 |= dupy[u] & ((c[u].w[3] != 0.0f) | (c[u].w[7] != 0.0f)); }
                const bool alias = (k >= 4 && dupz[u]) || (k == 2 && dupy[u]);        // this edge is the twin of an earlier one
                const bool paired = (i0 ^ i1) <= kGroupMask;
                const bool ld_wide = on & !alias & (n0 | (paired & n1)), ld_one = on & !alias & !paired & n1;
                wide_t w = wide_t();
                if (ld_wide) w = gather_ld<wide_t>(base + (i0 & ~kGroupMask));
                uint32_t s = 0u;
                if (ld_one) s = gather_ld<uint32_t>(base + i1);
                q[u][j] = w; a1[u][j] = s;
            }
        }
        // ---- interpolation (grid.h:144-163: fp16 fma over the corners in index order)
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            __half2 v[8];
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const int j = k >> 1;
                const uint32_t i0 = c[u].idx[k], i1 = c[u].idx[k + 1];
                const bool paired = (i0 ^ i1) <= kGroupMask;
#if NRC_GATHER_BITS == 128
                const uint32_t b0 = (i0 & 2u) ? ((i0 & 1u) ? q[u][j].w : q[u][j].z) : ((i0 & 1u) ? q[u][j].y : q[u][j].x);
                const uint32_t b1 = (i1 & 2u) ? ((i1 & 1u) ? q[u][j].w : q[u][j].z) : ((i1 & 1u) ? q[u][j].y : q[u][j].x);
#else
                const uint32_t b0 = (i0 & 1u) ? q[u][j].y : q[u][j].x;
                const uint32_t b1 = (i1 & 1u) ? q[u][j].y : q[u][j].x;
#endif
                const uint32_t b1s = paired ? b1 : a1[u][j];
                v[k] = *reinterpret_cast<const __half2*>(&b0);
                v[k + 1] = *reinterpret_cast<const __half2*>(&b1s);
            }
            if (dupy[u]) { v[2] = v[0]; v[3] = v[1]; }
            if (dupz[u]) { v[4] = v[0]; v[5] = v[1]; v[6] = v[2]; v[7] = v[3]; }
            __half2 r = __float2half2_rn(0.0f);
#pragma unroll
            for (int k = 0; k < 8; k++) r = __hfma2(__half2half2(__float2half_rn(c[u].w[k])), v[k], r);
            if (l0 + u < l_end) put.put2(2 * (l0 + u), r);
        }
    }
}

// ---- software-pipelined flavour (power-of-two tables, i.e. every preset): the gathers of level l+1 are in flight while level l
// is interpolated and the addresses of level l+2 are computed, so a warp overlaps its own address arithmetic (~150 instructions
// per level) with its own L2 round trips -- there are only four warps per scheduler to hide them otherwise (TMEM caps the
// resident tiles).  Levels are processed in runs of one KIND so that the index arithmetic carries no per-level selects:
// 1 = dense (x + y*s1 + z*s2), 2 = hashed (x ^ y*P1 ^ z*P2), 0 = generic (strides wrapped in uint32: aliasing corners).
struct LevelLoads {
    float f[3];          // fractional cell position
    uint32_t sel;        // bit j: low index bit of corner 2j, bit 4+j: of corner 2j+1, bit 8+j: both in one 8-byte word, bits 12/13: dupz/dupy
    uint2 q[4];          // aligned 8-byte word of edge j (corners 2j, 2j+1)
    uint32_t a1[4];      // corner 2j+1 when it lives elsewhere
};

template <int KIND>
__device__ __forceinline__ void level_issue(const EncParams& e, const __half2* __restrict__ grid, int l, float x0, float x1, float x2, LevelLoads& s) {
    const float scale = e.level_scale[l];
    float p[3] = {fmaf(scale, x0, 0.5f), fmaf(scale, x1, 0.5f), fmaf(scale, x2, 0.5f)};
    uint32_t g[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float t = floorf(p[d]);
        g[d] = (uint32_t)(int)t;
        s.f[d] = p[d] - t;
    }
    const uint32_t m = e.level_hsize[l] - 1u;
    const __half2* base = grid + e.level_offset[l];
    uint32_t h[4];       // index contribution of the four (y, z) corner combinations
    bool dupz = false, dupy = false;
    if (KIND == 2) {
        const uint32_t hy0 = g[1] * 2654435761u, hy1 = hy0 + 2654435761u, hz0 = g[2] * 805459861u, hz1 = hz0 + 805459861u;
        h[0] = hy0 ^ hz0; h[1] = hy1 ^ hz0; h[2] = hy0 ^ hz1; h[3] = hy1 ^ hz1;
    } else if (KIND == 1) {
        const uint32_t s1 = e.level_s1[l], s2 = e.level_s2[l], a = g[1] * s1 + g[2] * s2;
        h[0] = a; h[1] = a + s1; h[2] = a + s2; h[3] = a + s1 + s2;
    } else {
        const bool hashed = e.level_hash[l] != 0;
        const uint32_t s1 = e.level_s1[l], s2 = e.level_s2[l], a = g[1] * s1 + g[2] * s2;
        const uint32_t hy0 = g[1] * 2654435761u, hy1 = hy0 + 2654435761u, hz0 = g[2] * 805459861u, hz1 = hz0 + 805459861u;
        h[0] = hashed ? hy0 ^ hz0 : a; h[1] = hashed ? hy1 ^ hz0 : a + s1; h[2] = hashed ? hy0 ^ hz1 : a + s2; h[3] = hashed ? hy1 ^ hz1 : a + s1 + s2;
        dupz = !hashed && (s2 & m) == 0;
        dupy = dupz && (s1 & m) == 0;
    }
    const bool hashed_x = KIND == 2 || (KIND == 0 && e.level_hash[l] != 0);
    // a corner whose weight is exactly 0 is not loaded: the weight is a product of three factors f or 1 - f
    const bool zx[2] = {1.0f - s.f[0] != 0.0f, s.f[0] != 0.0f}, zy[2] = {1.0f - s.f[1] != 0.0f, s.f[1] != 0.0f}, zz[2] = {1.0f - s.f[2] != 0.0f, s.f[2] != 0.0f};
    uint32_t sel = (dupz ? 1u << 12 : 0u) | (dupy ? 1u << 13 : 0u);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t i0 = (hashed_x ? (g[0] ^ h[j]) : (g[0] + h[j])) & m, i1 = (hashed_x ? ((g[0] + 1u) ^ h[j]) : (g[0] + 1u + h[j])) & m;
        bool nyz = zy[j & 1] & zz[j >> 1];
        if (KIND == 0) {                                                     // twins of this edge that alias onto it
            if (j < 2) nyz |= dupz & zy[j & 1] & zz[1];
            if (j == 0) nyz |= dupy & (zy[1] & (zz[0] | zz[1]));
        }
        const bool alias = KIND == 0 && ((j >= 2 && dupz) || (j == 1 && dupy));
        const bool n0 = zx[0] & nyz, n1 = zx[1] & nyz;
        const bool paired = (i0 ^ i1) <= 1u;
        sel |= (i0 & 1u) << j | (i1 & 1u) << (4 + j) | (paired ? 1u : 0u) << (8 + j);
        uint2 w = make_uint2(0u, 0u);
        if (!alias & (n0 | (paired & n1))) w = gather_ld<uint2>(base + (i0 & ~1u));
        uint32_t o = 0u;
        if (!alias & !paired & n1) o = gather_ld<uint32_t>(base + i1);
        s.q[j] = w; s.a1[j] = o;
    }
    s.sel = sel;
}

// grid.h:144-163: fp16 fma over the eight corners in index order, weights computed in fp32 as (fx * fy) * fz
template <int KIND>
__device__ __forceinline__ __half2 level_finish(const LevelLoads& s) {
    __half2 v[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t b0 = (s.sel >> j & 1u) ? s.q[j].y : s.q[j].x;
        const uint32_t b1 = (s.sel >> (4 + j) & 1u) ? s.q[j].y : s.q[j].x;
        const uint32_t b1s = (s.sel >> (8 + j) & 1u) ? b1 : s.a1[j];
        v[2 * j] = *reinterpret_cast<const __half2*>(&b0);
        v[2 * j + 1] = *reinterpret_cast<const __half2*>(&b1s);
    }
    if (KIND == 0) {
        if (s.sel >> 13 & 1u) { v[2] = v[0]; v[3] = v[1]; }
        if (s.sel >> 12 & 1u) { v[4] = v[0]; v[5] = v[1]; v[6] = v[2]; v[7] = v[3]; }
    }
    const float fx[2] = {1.0f - s.f[0], s.f[0]}, fy[2] = {1.0f - s.f[1], s.f[1]}, fz[2] = {1.0f - s.f[2], s.f[2]};
    __half2 r = __float2half2_rn(0.0f);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float w = ((1.0f * fx[k & 1]) * fy[(k >> 1) & 1]) * fz[k >> 2];
        r = __hfma2(__half2half2(__float2half_rn(w)), v[k], r);
    }
    return r;
}

// one level per round, index arithmetic specialised per KIND (no pipelining: see the measurements at NRC_GATHER_PIPE)
template <int KIND, class Put>
__device__ __forceinline__ void hashgrid_run_simple(const EncParams& e, const __half2* __restrict__ grid, float x0, float x1, float x2, int lb, int le, Put& put) {
    for (int l = lb; l < le; l++) {
        LevelLoads A;
        level_issue<KIND>(e, grid, l, x0, x1, x2, A);
        put.put2(2 * l, level_finish<KIND>(A));
    }
}

template <int KIND, class Put>
__device__ __forceinline__ void hashgrid_run(const EncParams& e, const __half2* __restrict__ grid, float x0, float x1, float x2, int lb, int le, Put& put) {
    LevelLoads A, B;
    level_issue<KIND>(e, grid, lb, x0, x1, x2, A);
    for (int l = lb; l < le; l += 2) {
        const bool has_b = l + 1 < le;
        if (has_b) level_issue<KIND>(e, grid, l + 1, x0, x1, x2, B);
        put.put2(2 * l, level_finish<KIND>(A));
        if (has_b) {
            if (l + 2 < le) level_issue<KIND>(e, grid, l + 2, x0, x1, x2, A);
            put.put2(2 * (l + 1), level_finish<KIND>(B));
        }
    }
}

template <class Put>
__device__ __forceinline__ void hashgrid_pipelined(const EncParams& e, const __half2* __restrict__ grid, float x0, float x1, float x2,
                                                   int l_begin, int l_end, Put& put) {
    int l = l_begin;
    while (l < l_end) {
        const uint32_t kind = e.level_kind[l];
        int r = l + 1;
        while (r < l_end && e.level_kind[r] == kind) r++;
        if (NRC_GATHER_PIPE == 2) {
            if (kind == 2) hashgrid_run_simple<2>(e, grid, x0, x1, x2, l, r, put);
            else if (kind == 1) hashgrid_run_simple<1>(e, grid, x0, x1, x2, l, r, put);
            else hashgrid_run_simple<0>(e, grid, x0, x1, x2, l, r, put);
        } else {
            if (kind == 2) hashgrid_run<2>(e, grid, x0, x1, x2, l, r, put);
            else if (kind == 1) hashgrid_run<1>(e, grid, x0, x1, x2, l, r, put);
            else hashgrid_run<0>(e, grid, x0, x1, x2, l, r, put);
        }
        l = r;
    }
}

// put(k, half) / put2(k_even, half2) receive feature k of this record.
// Position features: hash-grid levels [l_begin, l_end) -- the whole encoding for the other (parameter-free) encoders.
// UNROLL > 0: rounds of UNROLL levels; UNROLL == 0: the software-pipelined flavour (needs ~128 registers: inference kernel only)
template <int UNROLL = 2, class Put>
__device__ __forceinline__ void encode_position(const EncParams& e, const __half2* __restrict__ grid, float x0, float x1, float x2,
                                                int l_begin, int l_end, Put& put, const GridAdamState* pf_state = nullptr) {
    if (e.pos_enc == POS_HASHGRID) {
        constexpr int U = UNROLL > 0 ? UNROLL : 2;
        if (e.all_pow2) {
            if (UNROLL == 0) hashgrid_pipelined(e, grid, x0, x1, x2, l_begin, l_end, put);
            else hashgrid_levels<U, true>(e, grid, x0, x1, x2, l_begin, l_end, put, pf_state);
        } else hashgrid_levels<U, false>(e, grid, x0, x1, x2, l_begin, l_end, put, pf_state);
    } else if (e.pos_enc == POS_IDENTITY) {
        put.put(0, __float2half_rn(x0)); put.put(1, __float2half_rn(x1)); put.put(2, __float2half_rn(x2));
    } else if (e.pos_enc == POS_TRIANGLE) {
        const float xs[3] = {x0, x1, x2};
#pragma unroll
        for (int d = 0; d < 3; d++)
            for (int f = 0; f < e.n_freq_pos; f++) {
                const float x = scalbnf(xs[d], f - 1);
                const float val = x + (float)f * 0.25f;
                put.put(d * e.n_freq_pos + f, __float2half_rn(fabsf(val - floorf(val) - 0.5f) * 4 - 1));
            }
    } else {
        const float xs[3] = {x0, x1, x2};
        const float PI = 3.14159265358979323846f;
#pragma unroll
        for (int d = 0; d < 3; d++)
            for (int f = 0; f < e.n_freq_pos; f++) {
                const float x = scalbnf(xs[d], f) * PI;
                put.put((d * e.n_freq_pos + f) * 2 + 0, __float2half_rn(__sinf(x)));
                put.put((d * e.n_freq_pos + f) * 2 + 1, __float2half_rn(__sinf(x + PI / 2)));
            }
    }
}

// Direction features and the padding columns (one thread does both: the padding overwrites direction columns under Q6).
template <class Put>
__device__ __forceinline__ void encode_direction_pad(const EncParams& e, float th, float ph, Put& put) {
    const __half one = __float2half_rn(1.0f);
    const int o = e.dir_off;
    const float ds[2] = {th, ph};
    if (e.dir_enc == DIR_ONEBLOB) {
        const float nb = (float)e.n_bins;
#pragma unroll
        for (int d = 0; d < 2; d++) {
            const float x = ds[d];
            if (e.oneblob_soa) {                                                               // oneblob.h:99-127
                float left = quartic_cdf(-x, nb) + quartic_cdf(-x - 1.0f, nb) + quartic_cdf(-x + 1.0f, nb);
                for (int k = 0; k < e.n_bins; k++) {
                    const float rb = (float)(k + 1) / nb;
                    const float right = quartic_cdf(rb - x, nb) + quartic_cdf(rb - x - 1.0f, nb) + quartic_cdf(rb - x + 1.0f, nb);
                    put.put(o + d * e.n_bins + k, __float2half_rn(right - left));
                    left = right;
                }
            } else {                                                                           // oneblob.h:46-70, 85-97
                const float first = quartic_cdf(-x, nb) + quartic_cdf(-x - 1.0f, nb) + quartic_cdf(-x + 1.0f, nb);
                float left = first;
                for (int k = 0; k < e.n_bins; k++) {
                    float right;
                    if (k == e.n_bins - 1) right = first + 1;
                    else { const float rb = (float)(k + 1) / nb; right = quartic_cdf(rb - x, nb) + quartic_cdf(rb - x - 1.0f, nb) + quartic_cdf(rb - x + 1.0f, nb); }
                    put.put(o + d * e.n_bins + k, __float2half_rn(right - left));
                    left = right;
                }
            }
        }
    } else if (e.dir_enc == DIR_IDENTITY) {
        put.put(o + 0, __float2half_rn(th)); put.put(o + 1, __float2half_rn(ph));
    } else {
#pragma unroll
        for (int d = 0; d < 2; d++)
            for (int f = 0; f < e.n_freq_dir; f++) {
                const float x = scalbnf(ds[d], f - 1);
                const float val = x + (float)f * 0.25f;
                put.put(o + d * e.n_freq_dir + f, __float2half_rn(fabsf(val - floorf(val) - 0.5f) * 4 - 1));
            }
    }
    // ---- padding (composite.h:137-218 pads with 1; SoA OneBlob writes the pad to the wrong rows, Q6)
    const int first_pad = o + e.dir_w;
    if (e.soa_bug) {
        const int n_pad = e.in_w - first_pad;
        for (int k = first_pad; k < e.in_w; k++) put.put(k, __float2half_rn(0.0f));
        for (int k = o + 2; k < o + 2 + n_pad; k++) put.put(k, one);
    } else {
        for (int k = first_pad; k < e.in_w; k++) put.put(k, one);
    }
}

template <int UNROLL = 2, class Put>
__device__ __forceinline__ void encode_record(const EncParams& e, const __half2* __restrict__ grid, float x0, float x1, float x2,
                                              float th, float ph, Put& put) {
    encode_position<UNROLL>(e, grid, x0, x1, x2, 0, e.n_levels, put);
    encode_direction_pad(e, th, ph, put);
}

// K-major canonical (no swizzle) tile in shared memory: this thread's row
struct SmemRowPut {
    uint8_t* row;   // tile + (r/8)*SBO + (r%8)*16
    __device__ __forceinline__ void put(int k, __half v) { *reinterpret_cast<__half*>(row + (k >> 3) * 128 + (k & 7) * 2) = v; }
    __device__ __forceinline__ void put2(int k, __half2 v) { *reinterpret_cast<__half2*>(row + (k >> 3) * 128 + (k & 7) * 2) = v; }
};
struct GlobalRowPut {
    __half* row;
    __device__ __forceinline__ void put(int k, __half v) { row[k] = v; }
    __device__ __forceinline__ void put2(int k, __half2 v) { *reinterpret_cast<__half2*>(row + k) = v; }
};

__global__ void __launch_bounds__(128) nrc_encode_kernel(const __grid_constant__ EncParams e, const __half* __restrict__ params, uint32_t n_mlp,
                                                         const float* __restrict__ in, uint32_t n, __half* __restrict__ out) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const float* rec = in + 5 * (size_t)i;
    GlobalRowPut put{out + (size_t)i * e.in_w};
    encode_record(e, reinterpret_cast<const __half2*>(params + n_mlp), rec[0], rec[1], rec[2], rec[3], rec[4], put);
}

// ---------------------------------------------------------------------------------------------- helpers
// development aid (NRCHPM_TRAIN_PROF=1): device-side timeline of a training step, tl[2k] = earliest start, tl[2k+1] = latest end of kernel k
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// programmatic dependent launch (PTX griddepcontrol): `pdl_wait` returns once the kernel launched before this one on the stream has
// completed and its memory is visible (no-op when the launch did not carry cudaLaunchAttributeProgrammaticStreamSerialization);
// `pdl_trigger` lets the next kernel's CTAs be dispatched as soon as every CTA of this grid has called it or exited, so the dependent's
// launch latency and prologue (TMEM allocation, barrier init) run underneath this kernel's tail
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void timeline_begin(unsigned long long* tl, int k) { if (tl && threadIdx.x == 0) atomicMin(tl + 2 * k, global_ns()); }
__device__ __forceinline__ void timeline_end(unsigned long long* tl, int k) { if (tl && threadIdx.x == 0) atomicMax(tl + 2 * k + 1, global_ns()); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// row-major [rows][K] fp16 matrix -> K-major canonical tile (B operand of the forward MMAs)
__device__ __forceinline__ void copy_weights_kmajor(uint8_t* dst, const __half* __restrict__ src, int rows, int K, int tid, int nthreads) {
    const int k8n = K >> 3, chunks = rows * k8n;
    for (int c = tid; c < chunks; c += nthreads) {
        const int n = c / k8n, k8 = c - n * k8n;
        *reinterpret_cast<int4*>(dst + (n >> 3) * (k8n * 128) + k8 * 128 + (n & 7) * 16) = *reinterpret_cast<const int4*>(src + (size_t)n * K + k8 * 8);
    }
}
// row-major [O][I] fp16 matrix used as B[n=i][k=o] -> MN-major canonical tile (backward MMAs; no transpose needed)
__device__ __forceinline__ void copy_weights_mnmajor(uint8_t* dst, const __half* __restrict__ src, int O, int I, int tid, int nthreads) {
    const int i8n = I >> 3, chunks = O * i8n;
    const int sbo = (O >> 3) * 128;
    for (int c = tid; c < chunks; c += nthreads) {
        const int o = c / i8n, i8 = c - o * i8n;
        *reinterpret_cast<int4*>(dst + i8 * sbo + (o & 7) * 16 + (o >> 3) * 128) = *reinterpret_cast<const int4*>(src + (size_t)o * I + i8 * 8);
    }
}

// the same two layouts, copied with cp.async (the caller waits: cp_async_wait_all + fence.proxy.async before the first MMA)
__device__ __forceinline__ void copy_weights_kmajor_async(uint8_t* dst, const __half* __restrict__ src, int rows, int K, int tid, int nthreads) {
    const int k8n = K >> 3, chunks = rows * k8n;
    for (int c = tid; c < chunks; c += nthreads) {
        const int n = c / k8n, k8 = c - n * k8n;
        tc05::cp_async16(dst + (n >> 3) * (k8n * 128) + k8 * 128 + (n & 7) * 16, src + (size_t)n * K + k8 * 8);
    }
}
__device__ __forceinline__ void copy_weights_mnmajor_async(uint8_t* dst, const __half* __restrict__ src, int O, int I, int tid, int nthreads) {
    const int i8n = I >> 3, chunks = O * i8n;
    const int sbo = (O >> 3) * 128;
    for (int c = tid; c < chunks; c += nthreads) {
        const int o = c / i8n, i8 = c - o * i8n;
        tc05::cp_async16(dst + i8 * sbo + (o & 7) * 16 + (o >> 3) * 128, src + (size_t)o * I + i8 * 8);
    }
}
// fire-and-forget fp16x2 reduction.  (atomicAdd(__half2*) on a generic pointer compiles to ATOM with a predicate result plus
// shared / local fall-back paths, and every call then waits for the round trip to L2.)
__device__ __forceinline__ void red_add_f16x2(__half2* addr, __half2 v) {
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(__cvta_generic_to_global(addr)), "r"(*reinterpret_cast<const uint32_t*>(&v)) : "memory");
}

// ---------------------------------------------------------------------------------------------- forward
struct FwdArgs {
    EncParams enc;
    const __half* params;       // fp16 parameter vector [network | encoding]: EMA weights for Inference(), working weights for training
    uint32_t n_mlp;
    int n_hidden;
    const float* in;            // records float[*][5]
    const uint32_t* indices;    // optional compaction list
    const uint32_t* d_count;    // optional device-side record count (<= n)
    uint32_t n;
    float* out;                 // inference: float[*][3]
    uint32_t smem_levels;       // inference: the first `smem_levels` hash-grid levels (`smem_level_entries` entries) are staged in smem
    uint32_t smem_level_entries;
    uint32_t ring_slots;        // nrc_wide_infer_ws_kernel: X tiles in the shared-memory ring
    // training only
    const float* target;        // float[n][3]
    __half* x16;                // [n][IN_W]   network input
    __half* acts;               // [H][n][64]  post-ReLU activations
    __half* out16;              // [n][16]
    __half* dout16;             // [n][16]     dL/doutput * loss_scale
    float* loss_partials;       // [n/128]
    float loss_scale;
    unsigned long long* tl;     // optional device-side timeline slot (development aid)
};

constexpr uint32_t kColD = 0, kColA = 96, kColsPerWg = 128;

// Shape of the inference instantiation (compile-time experiment knobs; defaults = measured best): warpgroups per CTA, CTAs per
// SM, hash-grid levels whose gathers are issued back to back.  More than 4 warpgroups per CTA pack their TMEM columns
// (64 accumulator + 32 operand columns each) at a stride of 96 instead of 128.
#ifndef NRC_INFER_WGS
#define NRC_INFER_WGS 2
#endif
#ifndef NRC_INFER_CTAS
#define NRC_INFER_CTAS 2
#endif
#ifndef NRC_INFER_UNROLL
#define NRC_INFER_UNROLL (NRC_GATHER_PIPE ? 0 : 1)      // 0 selects hashgrid_pipelined (PIPE 1: pipelined, PIPE 2: per-kind, one level per round)
#endif
#ifndef NRC_SMEM_LEVELS
#define NRC_SMEM_LEVELS 0            // coarse hash-grid levels staged in shared memory by the inference kernel
#endif
constexpr int kInferWgs = NRC_INFER_WGS, kInferCtas = NRC_INFER_CTAS;
__host__ __device__ constexpr uint32_t fwd_wg_stride(int wgs) { return wgs * 128 <= 512 ? 128u : 96u; }
__host__ __device__ constexpr uint32_t fwd_tmem_cols(int wgs) { return wgs * fwd_wg_stride(wgs) <= 128 ? 128u : wgs * fwd_wg_stride(wgs) <= 256 ? 256u : 512u; }
static_assert(kInferWgs >= 1 && kInferWgs <= 5 && kInferCtas * fwd_tmem_cols(kInferWgs) <= 512, "TMEM budget");

template <int IN_W>
__host__ __device__ constexpr size_t fwd_smem_bytes(int n_hidden, int wgs = 2) {
    return (size_t)IN_W * 128 + (size_t)(n_hidden - 1) * 8192 + 2048 + (size_t)wgs * IN_W * 256;
}

template <int IN_W, bool TRAIN>
__global__ void __launch_bounds__(TRAIN ? kFwdThreads : kInferWgs * 128, TRAIN ? 2 : kInferCtas) nrc_forward_kernel(const __grid_constant__ FwdArgs a) {
    using namespace tc05;
    constexpr int kMaxWg = TRAIN ? 2 : kInferWgs;
    constexpr uint32_t kStride = fwd_wg_stride(kMaxWg), kColAq = kStride - 32, kAlloc = fwd_tmem_cols(kMaxWg);
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[kMaxWg];
    __shared__ uint32_t tmem_base_s;
    __shared__ float loss_red[kMaxWg][4];
    const int tid = threadIdx.x, wg = tid >> 7, r = tid & 127, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x, nwg = nthreads >> 7;      // 2 warpgroups per CTA for large batches, 1 for small ones
    const int H = a.n_hidden;
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * 128;
    uint8_t* wo_s = wh_s + (H - 1) * 8192;
    uint8_t* x_s = wo_s + 2048 + wg * (IN_W * 256);
    uint8_t* lvl_s = wo_s + 2048 + nwg * (IN_W * 256);          // staged coarse hash-grid levels (inference only)

    if (warp == 0) { tmem_alloc(&tmem_base_s, kAlloc); tmem_relinquish(); }
    if (tid == 0) { for (int g = 0; g < kMaxWg; g++) mbar_init(&mbar[g], 1); fence_mbar_init(); }
    if (NRC_SMEM_LEVELS > 0 && !TRAIN && a.smem_levels) {
        const int4* src = reinterpret_cast<const int4*>(a.params + a.n_mlp);
        for (uint32_t c = tid; c < a.smem_level_entries / 4; c += nthreads) reinterpret_cast<int4*>(lvl_s)[c] = src[c];
    }
    copy_weights_kmajor(w0_s, a.params, kWidth, IN_W, tid, nthreads);
    for (int l = 1; l < H; l++) copy_weights_kmajor(wh_s + (l - 1) * 8192, a.params + IN_W * kWidth + (l - 1) * kWidth * kWidth, kWidth, kWidth, tid, nthreads);
    copy_weights_kmajor(wo_s, a.params + IN_W * kWidth + (H - 1) * kWidth * kWidth, kOutPad, kWidth, tid, nthreads);
    fence_proxy_async_smem();
    fence_before();
    __syncthreads();
    fence_after();

    const uint32_t tD = tmem_base_s + wg * kStride + kColD, tA = tmem_base_s + wg * kStride + kColAq;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc64 = make_idesc_f16(128, 64), idesc16 = make_idesc_f16(128, 16);
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
        if (NRC_SMEM_LEVELS > 0 && !TRAIN && a.smem_levels) {
            const uint32_t lvl_addr = smem_u32(lvl_s);
            for (uint32_t l = 0; l < a.smem_levels; l++) put.put2(2 * l, encode_level_smem(a.enc, l, lvl_addr + a.enc.level_offset[l] * 4u, x0, x1, x2));
            encode_position<NRC_INFER_UNROLL>(a.enc, grid, x0, x1, x2, (int)a.smem_levels, a.enc.n_levels, put);
            encode_direction_pad(a.enc, th, ph, put);
        } else {
            encode_record<TRAIN ? 2 : NRC_INFER_UNROLL>(a.enc, grid, x0, x1, x2, th, ph, put);
        }
        fence_proxy_async_smem();
        fence_before();
        named_bar_sync(1 + wg, 128);
        if (r == 0) {
            fence_after();
#pragma unroll
            for (int s = 0; s < IN_W / 16; s++)
                mma_f16_ss(tD, make_smem_desc(x_addr + s * 256, 128, IN_W * 16), make_smem_desc(w0_addr + s * 256, 128, IN_W * 16), idesc64, s > 0);
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
            // epilogue of hidden layer l: ReLU -> fp16 -> A operand of the next layer
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                uint32_t acc[32], p[16];
                tmem_ld32(tD + lane_base + hf * 32, acc);
                wait_ld();
                if (l == 0) {
                    // tcnn's ReLU is max(x, 0) in fp16, which maps NaN to 0 (NaN features reach layer 0 through the
                    // reference's phi = acos(>1)); cvt.relu would keep the NaN.  Later layers cannot see one.
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
                tmem_st16(tA + lane_base + hf * 16, p);
                if (TRAIN) {
                    int4* dst = reinterpret_cast<int4*>(a.acts + ((size_t)l * a.n + row) * kWidth + hf * 32);
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
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + l * 8192 + s * 256, 128, 1024), idesc64, s > 0);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wo_addr + s * 256, 128, 1024), idesc16, s > 0);
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
            float loss = 0;
            uint32_t g16[8];
#pragma unroll
            for (int j = 0; j < 8; j++) g16[j] = 0;
            float gk[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float diff = pr[k] - a.target[3 * (size_t)row + k];
                loss += diff * diff / denom / n_total;
                gk[k] = a.loss_scale * (2 * diff / denom) / n_total;
            }
            g16[0] = pack_f16x2(gk[0], gk[1]);
            g16[1] = pack_f16x2(gk[2], 0.0f);
            int4* gd = reinterpret_cast<int4*>(a.dout16 + (size_t)row * kOutPad);
            gd[0] = make_int4(g16[0], g16[1], 0, 0);
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
    if (warp == 0) tmem_dealloc(tmem_base_s, kAlloc);
}

// ---------------------------------------------------------------------------------------------- warp-specialised inference
// Round-2 inference kernel.  nrc_forward_kernel<.., false> gives every warpgroup a whole tile: 16 rounds of hash-grid gathers, then
// seven dependent tcgen05 layers; ncu (profiles/r01_ncu_fwd_infer_final.txt) showed the launch taking the SUM of the two phases
// (0.38 ms of gathers + 0.26 ms of MLP) because tensor memory caps the resident tiles at four per SM.  Here the two phases run in
// different warps of one persistent CTA per SM:
//   * NP producer warpgroups (no tensor memory, no accumulator registers): one record per thread, hash-grid gathers with the index
//     arithmetic specialised per level kind (dense / hashed / wrapped), OneBlob, -> a ring of NS K-major X tiles in shared memory;
//     `full[slot]` (128 arrivals) hands a tile over;
//   * NC consumer warpgroups: layer 0 straight from the ring slot (tcgen05.mma SS; its tcgen05.commit on `empty[slot]` returns the
//     slot as soon as the tensor core has read it), layers 1..H and the output layer as before (TS mode, activations stay in TMEM).
// The L1 -> L2 request path (the unit the gathers saturate, one 32-byte sector per request) is now fed continuously instead of in
// bursts between MLP phases.  Measured on B200, 2 073 600 random records, HashGrid16x2 + OneBlob4, 64 x 6 (profiles/r02_summary.md):
// tile-per-warpgroup kernel 0.633 ms; 4 producers + 2 consumers, 8 ring slots 0.576; 5 slots 0.562; 4 slots 0.539; 4 + 1, 4 slots
// 0.534 (default); 3 + 1, 3 slots 0.537; 5 + 1 0.566; 6 + 2 (64 registers, spills) 0.592; pipelined gathers: no change -- the fewer
// ring slots, the larger the L1 that serves the second corner of an x-edge, and 12 producer warps already saturate the request path.
// One consumer warpgroup sustains 0.36 ms per frame, enough next to the gathers but not for encodings without table look-ups:
// those keep the tile-per-warpgroup kernel (0.25 ms).
template <class Put>
__device__ __forceinline__ void hashgrid_by_kind(const EncParams& e, const __half2* __restrict__ grid, float x0, float x1, float x2, Put& put) {
    int l = 0;
    while (l < e.n_levels) {
        const uint32_t kind = e.level_kind[l];
        int r = l + 1;
        while (r < e.n_levels && e.level_kind[r] == kind) r++;
        if (kind == 2) hashgrid_run_simple<2>(e, grid, x0, x1, x2, l, r, put);
        else if (kind == 1) hashgrid_run_simple<1>(e, grid, x0, x1, x2, l, r, put);
        else hashgrid_run_simple<0>(e, grid, x0, x1, x2, l, r, put);
        l = r;
    }
}

// shape of the warp-specialised inference CTA (compile-time experiment knobs; defaults = measured best): producer / consumer
// warpgroups, ring slots for encoded widths up to 48 (wider tiles get proportionally fewer), pipelined gathers in the producers
#ifndef NRC_WS_NP
#define NRC_WS_NP 4
#endif
#ifndef NRC_WS_NC
#define NRC_WS_NC 1
#endif
#ifndef NRC_WS_SLOTS
#define NRC_WS_SLOTS 4
#endif
#ifndef NRC_WS_PIPE
#define NRC_WS_PIPE 0
#endif
#ifndef NRC_WS_STREAM
#define NRC_WS_STREAM 1
#endif
#if NRC_WS_STREAM
#define NRC_WS_LDCS(p) __ldcs(p)
#else
#define NRC_WS_LDCS(p) (*(p))
#endif
__host__ __device__ constexpr int ws_slots(int in_w) { return in_w <= 48 ? NRC_WS_SLOTS : (NRC_WS_SLOTS * 48) / in_w; }

template <int IN_W>
__host__ __device__ constexpr size_t infer_ws_smem_bytes(int n_hidden, int slots) {
    return (size_t)IN_W * 128 + (size_t)(n_hidden - 1) * 8192 + 2048 + (size_t)slots * IN_W * 256;
}

template <int IN_W, int NP, int NC, int NS>
// 80 registers (640 threads -> 51 200 of the SM's 65 536): a 256-thread CTA of the gradient-exchange kernel (48 registers) fits next to it,
// which is what lets a data-parallel replica run its inference underneath the exchanges (NrcCache::infer_and_train_overlapped)
__global__ void __maxnreg__(80) nrc_infer_ws_kernel(const __grid_constant__ FwdArgs a) {
    using namespace tc05;
    constexpr uint32_t kAlloc = NC <= 1 ? 128u : NC == 2 ? 256u : 512u;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[NS], empty[NS], mbar[NC];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, r = tid & 127;
    const int H = a.n_hidden;
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * 128;
    uint8_t* wo_s = wh_s + (H - 1) * 8192;
    uint8_t* ring = wo_s + 2048;

    timeline_begin(a.tl, 0);
    if (warp == 0) { tmem_alloc(&tmem_base_s, kAlloc); tmem_relinquish(); }
    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
        for (int c = 0; c < NC; c++) mbar_init(&mbar[c], 1);
        fence_mbar_init();
    }
    copy_weights_kmajor(w0_s, a.params, kWidth, IN_W, tid, (NP + NC) * 128);
    for (int l = 1; l < H; l++) copy_weights_kmajor(wh_s + (l - 1) * 8192, a.params + IN_W * kWidth + (l - 1) * kWidth * kWidth, kWidth, kWidth, tid, (NP + NC) * 128);
    copy_weights_kmajor(wo_s, a.params + IN_W * kWidth + (H - 1) * kWidth * kWidth, kOutPad, kWidth, tid, (NP + NC) * 128);
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
                // streaming loads: the records are read once and must not displace the hash-table lines from L1
                x0 = NRC_WS_LDCS(p); x1 = NRC_WS_LDCS(p + 1); x2 = NRC_WS_LDCS(p + 2); th = NRC_WS_LDCS(p + 3); ph = NRC_WS_LDCS(p + 4);
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
        const uint32_t tD = tmem_base_s + c * 128, tA = tD + 64;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t idesc64 = make_idesc_f16(128, 64), idesc16 = make_idesc_f16(128, 16);
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
                    mma_f16_ss(tD, make_smem_desc(x_addr + s * 256, 128, IN_W * 16), make_smem_desc(w0_addr + s * 256, 128, IN_W * 16), idesc64, s > 0);
                mma_commit(&empty[slot]);
                mma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
            for (int l = 0; l < H; l++) {
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    uint32_t acc[32], p[16];
                    tmem_ld32(tD + lane_base + hf * 32, acc);
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
                    tmem_st16(tA + lane_base + hf * 16, p);
                }
                wait_st();
                fence_before();
                named_bar_sync(1 + c, 128);
                if (r == 0) {
                    fence_after();
                    if (l < H - 1) {
#pragma unroll
                        for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + l * 8192 + s * 256, 128, 1024), idesc64, s > 0);
                    } else {
#pragma unroll
                        for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wo_addr + s * 256, 128, 1024), idesc16, s > 0);
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
                for (int q = 0; q < 3; q++) {
                    const float v = __half2float(__float2half_rn(__uint_as_float(o[q])));
                    if (NRC_WS_STREAM) __stcs(dst + q, v); else dst[q] = v;
                }
            }
            // (every thread's read of tD is complete -- wait::ld -- before the group's next barrier, which precedes the next MMA into tD ...
            fence_before();
            named_bar_sync(1 + c, 128);      // ... but the NEXT tile's layer-0 MMA is issued by thread 0 right after its wait on `full`: order it here)
        }
    }
    fence_before();
    __syncthreads();
    timeline_end(a.tl, 0);
    if (warp == 0) tmem_dealloc(tmem_base_s, kAlloc);
}

// ---------------------------------------------------------------------------------------------- backward
struct BwdArgs {
    EncParams enc;
    const __half* params;       // working weights
    uint32_t n_mlp;
    int n_hidden;
    uint32_t n;
    const float* in;            // records (hash-grid scatter recomputes the cells)
    const __half* acts;         // [H][n][64]
    const __half* dout16;       // [n][16]
    __half* dacts;              // [H][n][64] gradients w.r.t. the hidden pre-activations (ReLU mask applied)
    __half* dx16;               // [n][IN_W] or null
    __half* grid_grad;          // fp16 gradient of the encoding parameters or null
    const float* loss_partials;
    float* loss_out;
    uint32_t n_loss_partials;
    int need_dx;
};

template <int IN_W>
__host__ __device__ constexpr size_t bwd_smem_bytes(int n_hidden) {
    return (size_t)IN_W * 128 + (size_t)(n_hidden - 1) * 8192 + 2048;
}

template <int IN_W>
__global__ void __launch_bounds__(kFwdThreads, 2) nrc_backward_kernel(const __grid_constant__ BwdArgs a) {
    using namespace tc05;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, wg = tid >> 7, r = tid & 127, warp = tid >> 5;
    const int nthreads = blockDim.x, nwg = nthreads >> 7;
    const int H = a.n_hidden;
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * 128;
    uint8_t* wo_s = wh_s + (H - 1) * 8192;

    if (warp == 0) { tmem_alloc(&tmem_base_s, 2 * kColsPerWg); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_mbar_init(); }
    if (blockIdx.x == 0 && tid == 0 && a.loss_out) {     // Trainer::loss (trainer.h:205-207): fixed-order sum of the tile partials
        float s = 0;
        for (uint32_t i = 0; i < a.n_loss_partials; i++) s += a.loss_partials[i];
        *a.loss_out = s;
    }
    copy_weights_mnmajor(w0_s, a.params, kWidth, IN_W, tid, nthreads);
    for (int l = 1; l < H; l++) copy_weights_mnmajor(wh_s + (l - 1) * 8192, a.params + IN_W * kWidth + (l - 1) * kWidth * kWidth, kWidth, kWidth, tid, nthreads);
    copy_weights_mnmajor(wo_s, a.params + IN_W * kWidth + (H - 1) * kWidth * kWidth, kOutPad, kWidth, tid, nthreads);
    fence_proxy_async_smem();
    fence_before();
    __syncthreads();
    fence_after();

    const uint32_t tD = tmem_base_s + wg * kColsPerWg + kColD, tA = tmem_base_s + wg * kColsPerWg + kColA;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc64 = make_idesc_f16(128, 64, 0, 1), idescx = make_idesc_f16(128, IN_W, 0, 1);
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
            mma_f16_ts(tD, tA, make_smem_desc(wo_addr, 128, 256), idesc64, 0);
            mma_commit(bar);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();

        for (int l = H - 1; l >= 0; l--) {
            const int4* ap = reinterpret_cast<const int4*>(a.acts + ((size_t)l * a.n + row) * kWidth);
            int4* dp = reinterpret_cast<int4*>(a.dacts + ((size_t)l * a.n + row) * kWidth);
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                uint32_t acc[32], p[16];
                int4 av[4];
#pragma unroll
                for (int c = 0; c < 4; c++) av[c] = ap[hf * 4 + c];
                tmem_ld32(tD + lane_base + hf * 32, acc);
                wait_ld();
                const uint32_t* aw = reinterpret_cast<const uint32_t*>(av);
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const float lo = (aw[j] & 0x0000ffffu) ? __uint_as_float(acc[2 * j]) : 0.0f;
                    const float hi = (aw[j] & 0xffff0000u) ? __uint_as_float(acc[2 * j + 1]) : 0.0f;
                    p[j] = pack_f16x2(lo, hi);
                }
                tmem_st16(tA + lane_base + hf * 16, p);
#pragma unroll
                for (int c = 0; c < 4; c++) dp[hf * 4 + c] = make_int4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
            }
            if (l == 0 && !a.need_dx) break;
            wait_st();
            fence_before();
            named_bar_sync(1 + wg, 128);
            if (r == 0) {
                fence_after();
                if (l > 0) {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + (l - 1) * 8192 + s * 256, 128, 1024), idesc64, s > 0);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(w0_addr + s * 256, 128, 1024), idescx, s > 0);
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
                for (int j = 0; j < 16; j++) dst[j] = pack_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
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
                // kernel_grid_backward (grid.h:215-320): (half)weight * grad, fp16x2 atomics
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
    if (warp == 0) tmem_dealloc(tmem_base_s, 2 * kColsPerWg);
}

// ---------------------------------------------------------------------------------------------- two threads per record
// Second generation of the forward / backward kernels: a tile of 128 records is owned by a GROUP of 256 threads.  Warps w and
// w + 4 of a group share one TMEM lane quadrant (= the same 32 records) and split the per-record work: hash-grid levels
// [0, L/2) vs [L/2, L) + direction encoding, accumulator columns [0, 32) vs [32, 64) in every epilogue, and the gradient
// scatter by levels again.  The dependent chain per tile (gather rounds, tcgen05.ld -> cvt -> tcgen05.st per layer, atomics)
// is half as long and twice as many warps are in flight per tile -- what the latency-bound 128-tile training batch needs.
// A CTA carries blockDim.x / 256 groups (tiles in flight); one CTA per SM.
constexpr int kGroupThreads = 256;

template <int IN_W>
__host__ __device__ constexpr size_t fwd2_smem_bytes(int n_hidden, int groups) {
    return (size_t)IN_W * 128 + (size_t)(n_hidden - 1) * 8192 + 2048 + (size_t)groups * IN_W * 256;
}
__host__ __device__ constexpr uint32_t tmem_cols_for_groups(int groups) { return groups <= 1 ? 128u : groups == 2 ? 256u : 512u; }

template <int IN_W, bool TRAIN>
__global__ void __launch_bounds__(TRAIN ? 256 : 768, TRAIN ? 2 : 1) nrc_forward2_kernel(const __grid_constant__ FwdArgs a) {
    using namespace tc05;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[4], wbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float loss_red[4][4];
    const int tid = threadIdx.x, grp = tid >> 8, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, half = (warp >> 2) & 1, r = quad * 32 + lane;
    const int nthreads = blockDim.x, ngrp = nthreads >> 8;
    const int H = a.n_hidden;
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * 128;
    uint8_t* wo_s = wh_s + (H - 1) * 8192;
    uint8_t* x_s = wo_s + 2048 + grp * (IN_W * 256);

    if (warp == 0) { tmem_alloc(&tmem_base_s, tmem_cols_for_groups(ngrp)); tmem_relinquish(); }
    if (tid == 0) { for (int g = 0; g < ngrp; g++) mbar_init(&mbar[g], 1); mbar_init(&wbar, nthreads); fence_mbar_init(); }
    // the weights travel to shared memory asynchronously while the first tile is encoded; every thread arrives on `wbar` once its
    // copies have landed (and are visible to the tensor core's proxy), the group leaders wait on it before their first MMA
    copy_weights_kmajor_async(w0_s, a.params, kWidth, IN_W, tid, nthreads);
    for (int l = 1; l < H; l++) copy_weights_kmajor_async(wh_s + (l - 1) * 8192, a.params + IN_W * kWidth + (l - 1) * kWidth * kWidth, kWidth, kWidth, tid, nthreads);
    copy_weights_kmajor_async(wo_s, a.params + IN_W * kWidth + (H - 1) * kWidth * kWidth, kOutPad, kWidth, tid, nthreads);
    bool weights_pending = true;
    fence_before();
    __syncthreads();
    fence_after();

    const uint32_t tD = tmem_base_s + grp * kColsPerWg + kColD, tA = tmem_base_s + grp * kColsPerWg + kColA;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t idesc64 = make_idesc_f16(128, 64), idesc16 = make_idesc_f16(128, 16);
    const uint32_t x_addr = smem_u32(x_s), w0_addr = smem_u32(w0_s), wh_addr = smem_u32(wh_s), wo_addr = smem_u32(wo_s);
    const __half2* grid = reinterpret_cast<const __half2*>(a.params + a.n_mlp);
    uint64_t* bar = &mbar[grp];
    uint32_t phase = 0;
    const bool leader = (tid & 255) == 0;
    const bool hashgrid = a.enc.pos_enc == POS_HASHGRID;
    const int l_split = hashgrid ? (a.enc.n_levels + 1) / 2 : a.enc.n_levels;

    uint32_t n = a.n;
    if (a.d_count) n = min(n, *a.d_count);
    const uint32_t n_tiles = (n + kTile - 1) / kTile;

    for (uint32_t tile = blockIdx.x * ngrp + grp; tile < n_tiles; tile += gridDim.x * ngrp) {
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
        {   // one call site (code size: each CTA runs this once in the 128-tile training batch, instruction fetch is cold)
            const int lb = half == 0 ? 0 : l_split, le = (half == 0 || !hashgrid) ? l_split : a.enc.n_levels;
            if (half == 0 || hashgrid) encode_position<TRAIN ? NRC_TRAIN_UNROLL : 2>(a.enc, grid, x0, x1, x2, lb, le, put);
            if (half == 1) encode_direction_pad(a.enc, th, ph, put);
        }
        if (weights_pending) { cp_async_wait_all(); fence_proxy_async_smem(); mbar_arrive(&wbar); }
        fence_proxy_async_smem();
        fence_before();
        named_bar_sync(1 + grp, kGroupThreads);
        if (leader) {
            if (weights_pending) mbar_wait(&wbar, 0);
            fence_after();
#pragma unroll
            for (int s = 0; s < IN_W / 16; s++)
                mma_f16_ss(tD, make_smem_desc(x_addr + s * 256, 128, IN_W * 16), make_smem_desc(w0_addr + s * 256, 128, IN_W * 16), idesc64, s > 0);
            mma_commit(bar);
        }
        weights_pending = false;
        if (TRAIN) {
            int4* dst = reinterpret_cast<int4*>(a.x16 + (size_t)row * IN_W);
#pragma unroll
            for (int c = 0; c < IN_W / 16; c++) dst[half * (IN_W / 16) + c] = *reinterpret_cast<const int4*>(my_row + (half * (IN_W / 16) + c) * 128);
        }
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();

        for (int l = 0; l < H; l++) {
            // epilogue of hidden layer l, this thread's 32 columns: ReLU -> fp16 -> A operand of the next layer
            uint32_t acc[32], p[16];
            tmem_ld32(tD + lane_base + half * 32, acc);
            wait_ld();
            if (l == 0) {
                // tcnn's ReLU is max(x, 0) in fp16, which maps NaN to 0 (NaN features reach layer 0 through the
                // reference's phi = acos(>1)); cvt.relu would keep the NaN.  Later layers cannot see one.
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
            tmem_st16(tA + lane_base + half * 16, p);
            if (TRAIN) {
                int4* dst = reinterpret_cast<int4*>(a.acts + ((size_t)l * a.n + row) * kWidth + half * 32);
#pragma unroll
                for (int c = 0; c < 4; c++) dst[c] = make_int4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
            }
            wait_st();
            fence_before();
            named_bar_sync(1 + grp, kGroupThreads);
            if (leader) {
                fence_after();
                if (l < H - 1) {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + l * 8192 + s * 256, 128, 1024), idesc64, s > 0);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wo_addr + s * 256, 128, 1024), idesc16, s > 0);
                }
                mma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
        }
        // output layer (16 accumulator columns): fp16 like tcnn's network output, then float (common_device.h:990-999)
        float loss = 0;
        if (half == 0) {
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
                float gk[3];
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
                if (lane == 0) loss_red[grp][quad] = loss;
            }
        }
        if (TRAIN) {
            named_bar_sync(1 + grp, kGroupThreads);
            if (leader) a.loss_partials[tile] = ((loss_red[grp][0] + loss_red[grp][1]) + loss_red[grp][2]) + loss_red[grp][3];
        }
    }
    if (weights_pending) { cp_async_wait_all(); mbar_arrive(&wbar); }     // a group without tiles
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, tmem_cols_for_groups(ngrp));
}

template <int IN_W>
__host__ __device__ constexpr size_t bwd2_smem_bytes(int n_hidden) {
    return (size_t)IN_W * 128 + (size_t)(n_hidden - 1) * 8192 + 2048;
}

template <int IN_W>
__global__ void __launch_bounds__(256, 2) nrc_backward2_kernel(const __grid_constant__ BwdArgs a) {
    using namespace tc05;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[4], wbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, grp = tid >> 8, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, half = (warp >> 2) & 1, r = quad * 32 + lane;
    const int nthreads = blockDim.x, ngrp = nthreads >> 8;
    const int H = a.n_hidden;
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * 128;
    uint8_t* wo_s = wh_s + (H - 1) * 8192;

    if (warp == 0) { tmem_alloc(&tmem_base_s, tmem_cols_for_groups(ngrp)); tmem_relinquish(); }
    if (tid == 0) { for (int g = 0; g < ngrp; g++) mbar_init(&mbar[g], 1); mbar_init(&wbar, nthreads); fence_mbar_init(); }
    if (blockIdx.x == 0 && warp == 1 && a.loss_out) {     // Trainer::loss (trainer.h:205-207): fixed-order sum of the tile partials
        // lanes read in parallel, lane 0 adds in tile order (deterministic)
        float s = 0;
        for (uint32_t i0 = 0; i0 < a.n_loss_partials; i0 += 32) {
            const float v = (i0 + lane < a.n_loss_partials) ? a.loss_partials[i0 + lane] : 0.0f;
#pragma unroll
            for (int j = 0; j < 32; j++) { const float vj = __shfl_sync(0xffffffffu, v, j); if (i0 + j < a.n_loss_partials) s += vj; }
        }
        if (lane == 0) *a.loss_out = s;
    }
    copy_weights_mnmajor_async(wo_s, a.params + IN_W * kWidth + (H - 1) * kWidth * kWidth, kOutPad, kWidth, tid, nthreads);
    for (int l = H - 1; l >= 1; l--) copy_weights_mnmajor_async(wh_s + (l - 1) * 8192, a.params + IN_W * kWidth + (l - 1) * kWidth * kWidth, kWidth, kWidth, tid, nthreads);
    copy_weights_mnmajor_async(w0_s, a.params, kWidth, IN_W, tid, nthreads);
    bool weights_pending = true;
    fence_before();
    __syncthreads();
    fence_after();

    const uint32_t tD = tmem_base_s + grp * kColsPerWg + kColD, tA = tmem_base_s + grp * kColsPerWg + kColA;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t idesc64 = make_idesc_f16(128, 64, 0, 1), idescx = make_idesc_f16(128, IN_W, 0, 1);
    const uint32_t w0_addr = smem_u32(w0_s), wh_addr = smem_u32(wh_s), wo_addr = smem_u32(wo_s);
    uint64_t* bar = &mbar[grp];
    uint32_t phase = 0;
    const bool leader = (tid & 255) == 0;
    const uint32_t n_tiles = a.n / kTile;
    const bool scatter = a.grid_grad && a.enc.pos_enc == POS_HASHGRID;
    const int l_split = (a.enc.n_levels + 1) / 2;

    for (uint32_t tile = blockIdx.x * ngrp + grp; tile < n_tiles; tile += gridDim.x * ngrp) {
        const uint32_t row = tile * kTile + r;
        // requested early, consumed late: the record position (gradient scatter) and the last layer's activations (ReLU mask)
        float x0 = 0, x1 = 0, x2 = 0;
        if (scatter) { const float* rp = a.in + 5 * (size_t)row; x0 = rp[0]; x1 = rp[1]; x2 = rp[2]; }
        int4 av[4];
        {
            const int4* ap = reinterpret_cast<const int4*>(a.acts + ((size_t)(H - 1) * a.n + row) * kWidth + half * 32);
#pragma unroll
            for (int c = 0; c < 4; c++) av[c] = ap[c];
        }
        if (half == 0) {   // dL/doutput row -> A operand (K = 16)
            const int4* src = reinterpret_cast<const int4*>(a.dout16 + (size_t)row * kOutPad);
            const int4 v0 = src[0], v1 = src[1];
            uint32_t p[8] = {(uint32_t)v0.x, (uint32_t)v0.y, (uint32_t)v0.z, (uint32_t)v0.w, (uint32_t)v1.x, (uint32_t)v1.y, (uint32_t)v1.z, (uint32_t)v1.w};
            tmem_st8(tA + lane_base, p);
        }
        wait_st();
        if (weights_pending) { cp_async_wait_all(); fence_proxy_async_smem(); mbar_arrive(&wbar); }
        fence_before();
        named_bar_sync(1 + grp, kGroupThreads);
        if (leader) {
            if (weights_pending) mbar_wait(&wbar, 0);
            fence_after();
            mma_f16_ts(tD, tA, make_smem_desc(wo_addr, 128, 256), idesc64, 0);
            mma_commit(bar);
        }
        weights_pending = false;
        mbar_wait(bar, phase); phase ^= 1;
        fence_after();

        for (int l = H - 1; l >= 0; l--) {
            uint32_t acc[32], p[16];
            tmem_ld32(tD + lane_base + half * 32, acc);
            wait_ld();
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(av);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const float lo = (aw[j] & 0x0000ffffu) ? __uint_as_float(acc[2 * j]) : 0.0f;
                const float hi = (aw[j] & 0xffff0000u) ? __uint_as_float(acc[2 * j + 1]) : 0.0f;
                p[j] = pack_f16x2(lo, hi);
            }
            if (l > 0) {   // next layer's mask: in flight while the MMA below runs
                const int4* ap = reinterpret_cast<const int4*>(a.acts + ((size_t)(l - 1) * a.n + row) * kWidth + half * 32);
#pragma unroll
                for (int c = 0; c < 4; c++) av[c] = ap[c];
            }
            tmem_st16(tA + lane_base + half * 16, p);
            int4* dp = reinterpret_cast<int4*>(a.dacts + ((size_t)l * a.n + row) * kWidth + half * 32);
#pragma unroll
            for (int c = 0; c < 4; c++) dp[c] = make_int4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
            if (l == 0 && !a.need_dx) break;
            wait_st();
            fence_before();
            named_bar_sync(1 + grp, kGroupThreads);
            if (leader) {
                fence_after();
                if (l > 0) {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + (l - 1) * 8192 + s * 256, 128, 1024), idesc64, s > 0);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(w0_addr + s * 256, 128, 1024), idescx, s > 0);
                }
                mma_commit(bar);
            }
            mbar_wait(bar, phase); phase ^= 1;
            fence_after();
        }
        if (a.need_dx) {
            if (a.dx16) {
                // dL/d(network input): fp16 like tcnn's fc_multiply output (fully_fused_mlp.cu:832-835); this thread's half of the row
                uint32_t* dst = reinterpret_cast<uint32_t*>(a.dx16 + (size_t)row * IN_W);
#pragma unroll
                for (int c = 0; c < IN_W / 2; c += 8) {
                    uint32_t t[8];
                    tmem_ld8(tD + lane_base + half * (IN_W / 2) + c, t);
                    wait_ld();
#pragma unroll
                    for (int j = 0; j < 4; j++) dst[(half * (IN_W / 2) + c) / 2 + j] = pack_f16x2(__uint_as_float(t[2 * j]), __uint_as_float(t[2 * j + 1]));
                }
            }
            if (scatter) {
                // kernel_grid_backward (grid.h:215-320): (half)weight * grad, fp16x2 atomics; levels split between the two halves.
                // The level loop stays rolled (the gradient pair of level l is read from TMEM columns 2l, 2l+1): unrolled 16 x 8
                // corners the kernel spent 40 % of its issue slots waiting for instruction fetch.
                __half2* gg = reinterpret_cast<__half2*>(a.grid_grad);
                const int lb = half == 0 ? 0 : l_split, le = half == 0 ? l_split : a.enc.n_levels;
#pragma unroll 1
                for (int l = lb; l < le; l++) {
                    uint32_t t[2];
                    tmem_ld2(tD + lane_base + 2 * l, t);
                    wait_ld();
                    const uint32_t gp = pack_f16x2(__uint_as_float(t[0]), __uint_as_float(t[1]));
                    const __half2 g = *reinterpret_cast<const __half2*>(&gp);
                    GridLevel c;
                    grid_level_cell(a.enc, l, x0, x1, x2, c);
                    __half2* base = gg + a.enc.level_offset[l];
#pragma unroll
                    for (int k = 0; k < 8; k++) red_add_f16x2(base + c.idx[k], __hmul2(__half2half2(__float2half_rn(c.w[k])), g));
                }
            }
            // (every thread's reads of tD are complete -- wait::ld -- before it reaches the next tile's first barrier)
        }
    }
    if (weights_pending) { cp_async_wait_all(); mbar_arrive(&wbar); }     // a group without tiles
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, tmem_cols_for_groups(ngrp));
}

// ---------------------------------------------------------------------------------------------- weight gradients
// dW_m = sum over the batch of dY_m^T * A_{m-1}: batch is the contraction dimension, both operands are "MN-major"
// ([sample][feature] rows copied as they are).  blockIdx.x = batch chunk, blockIdx.y = weight matrix.  Each CTA writes
// its fp32 partial; the optimizer kernel adds the chunks in a fixed order (deterministic, no atomics).
struct DwArgs {
    int n_hidden;
    uint32_t n, kc, n_mlp;
    const __half* x16;
    const __half* acts;
    const __half* dacts;
    const __half* dout16;
    float* partials;            // [n_chunks][n_mlp]
};

constexpr size_t kDwSmemBytes = 2 * (16384 + 20480);

template <int IN_W>
__global__ void __launch_bounds__(128, 1) nrc_dw_kernel(const __grid_constant__ DwArgs a) {
    using namespace tc05;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = a.n_hidden, m = blockIdx.y;
    // operands of matrix m: A-op (M = 64 rows of D), B-op (N columns of D)
    const __half* aop; const __half* bop; int aw, bw;
    if (m == 0) { aop = a.dacts; aw = kWidth; bop = a.x16; bw = IN_W; }
    else if (m < H) { aop = a.dacts + (size_t)m * a.n * kWidth; aw = kWidth; bop = a.acts + (size_t)(m - 1) * a.n * kWidth; bw = kWidth; }
    else { aop = a.acts + (size_t)(H - 1) * a.n * kWidth; aw = kWidth; bop = a.dout16; bw = kOutPad; }

    if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_mbar_init(); }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tD = tmem_base_s;
    const uint32_t idesc = make_idesc_f16(64, (uint32_t)bw, 1, 1);
    const uint32_t b0 = blockIdx.x * a.kc, b1 = min(a.n, b0 + a.kc);
    uint32_t phase[2] = {0, 0};
    int it = 0;
    for (uint32_t b = b0; b < b1; b += kTile, it++) {
        const int buf = it & 1;
        uint8_t* as = smem + buf * (16384 + 20480);
        uint8_t* bs = as + 16384;
        if (it >= 2) { mbar_wait(&mbar[buf], phase[buf]); phase[buf] ^= 1; }     // MMAs that read this buffer are done
        // thread = sample: 16-byte chunks land conflict-free at (mn8)*2048 + (b%8)*16 + (b/8)*128
        const uint32_t off = (tid & 7) * 16 + (tid >> 3) * 128;
        const int4* ag = reinterpret_cast<const int4*>(aop + (size_t)(b + tid) * aw);
#pragma unroll
        for (int c = 0; c < kWidth / 8; c++) *reinterpret_cast<int4*>(as + c * 2048 + off) = ag[c];
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
    // drain: the last commit covers every earlier MMA
    {
        const int last = (it - 1) & 1;
        mbar_wait(&mbar[last], phase[last]);
        fence_after();
    }
    // M = 64 accumulators: row 16*warp + lane (lane < 16) of D lives in TMEM lane 32*warp + lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int row = warp * 16 + lane;
    float* out = a.partials + (size_t)blockIdx.x * a.n_mlp;
    size_t moff = m == 0 ? 0 : (size_t)IN_W * kWidth + (size_t)(m - 1) * kWidth * kWidth;
    for (int c = 0; c < bw; c += 16) {
        uint32_t v[16];
        tmem_ld16(tD + lane_base + c, v);
        wait_ld();
        if (lane < 16) {
            if (m < H) {
                float4* dst = reinterpret_cast<float4*>(out + moff + (size_t)row * bw + c);
#pragma unroll
                for (int q = 0; q < 4; q++) dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            } else {
                // output layer computed transposed: D[i][o] -> W_out[o][i]
#pragma unroll
                for (int q = 0; q < 16; q++) out[moff + (size_t)(c + q) * kWidth + row] = __uint_as_float(v[q]);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 128);
}

// ---------------------------------------------------------------------------------------------- fused training step
// Forward, RelativeL2Luminance, backward, weight gradients and the hash-grid gradient scatter of ONE 128-record tile in ONE CTA
// (replaces nrc_forward2_kernel<TRAIN> + nrc_backward2_kernel + nrc_dw_kernel and the three activation tensors they exchanged
// through HBM).  Reference counterpart: Trainer::training_step -> forward / loss / backward (trainer.h:163-190), i.e.
// kernel_grid, kernel_mlp_fused, relative_l2_luminance_loss, kernel_mlp_fused_backward, the split-K weight-gradient GEMMs
// (fully_fused_mlp.cu:783-836) and kernel_grid_backward (grid.h:215-320).
//
//  * shared memory (192 KB for 64 x 6, 48 inputs): ONE image of the weights, the encoded tile X, the post-ReLU activations of
//    every hidden layer, two buffers for the pre-activation gradients dY, the loss gradient.  Every tile is stored
//    [record][feature] in the canonical no-swizzle core-matrix layout: 8 records x 8 features (16 bytes per record) per core
//    matrix.  The same bytes serve as K-major operand when `feature` is the contraction dimension (forward: A = X, B = W) and as
//    MN-major operand when `record` (dW = dY^T A: both operands) or `out` (backward: B = W read as [in][out]) is -- only the
//    LBO / SBO fields of the descriptor swap roles (LBO = stride between core matrices along K, SBO = along M / N).
//  * tensor memory (512 columns): 64 accumulator columns + 32 operand columns for the layer chain exactly as in the inference
//    kernel, and one fp32 accumulator region per weight matrix (M = 64) that collects dW over ALL tiles of the CTA; it is
//    written to the CTA's partial once, at the end (the optimizer adds the partials in a fixed order: deterministic).
//  * TPR threads per record (2 or 4): warps w, w + 4, ... share a TMEM lane quadrant and split the accumulator columns of every
//    epilogue, the hash-grid levels of the gathers and of the gradient scatter.
//  * MMA issue order per backward layer: the chain MMA (dA_{l-1} = dY_l W_l) first, committed to `bar_mma`; then the weight-
//    gradient MMAs of that layer, committed to `bar_dw[l & 1]`, which only guards the re-use of the dY buffer two layers later.
struct TrainArgs {
    EncParams enc;
    const __half* params;       // working weights [network | encoding]
    uint32_t n_mlp;
    int n_hidden;
    const float* in;            // records float[n][5]
    const float* target;        // float[n][3]
    uint32_t n;                 // multiple of 128
    float loss_scale;
    __half* grid_grad;          // fp16 gradient of the encoding parameters or null
    float* dw_partials;         // [gridDim.x][n_mlp]
    float* loss_partials;       // [n / 128]
    float* loss_out;
    unsigned int* done_counter; // zero before the first launch; the kernel leaves it zero
    __half* out16;              // [n][16] network output (fp16 like tcnn's), kept for nrc_last_step_tensor
    __half* dout16;             // [n][16] dL/doutput * loss_scale
    const GridAdamState* grid_state;   // optimizer state of the encoding entries: prefetched into L2 for the entries this step touches
    long long* prof;            // optional [gridDim.x][16] phase time stamps (clock64 of thread 0), development aid
    unsigned long long* tl;     // optional device-side timeline (development aid)
};

// dense EMA of the hash-grid part of the fp16 weight vector (ema.h:63-76): grid-stride loop over 16-byte words, four words of each
// array in flight per thread
__device__ __forceinline__ void grid_ema_pass(const __half* w16, __half* ema16, uint64_t n_vec, float decay, float debias_old, float debias_new,
                                              uint64_t first, uint64_t stride) {
    const int4* wv = reinterpret_cast<const int4*>(w16);
    int4* ev = reinterpret_cast<int4*>(ema16);
    for (uint64_t v0 = first; v0 < n_vec; v0 += 4 * stride) {
        union V8 { int4 v; __half h[8]; };
        V8 w[4], e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t v = v0 + u * stride;
            if (v < n_vec) { w[u].v = wv[v]; e[u].v = ev[v]; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t v = v0 + u * stride;
            if (v >= n_vec) break;
#pragma unroll
            for (int j = 0; j < 8; j++) e[u].h[j] = __float2half_rn((__half2float(e[u].h[j]) * decay * debias_old + __half2float(w[u].h[j]) * (1 - decay)) * debias_new);
            ev[v] = e[u].v;
        }
    }
}
#define NRC_PROF(k) do { if (a.prof && tid == 0) a.prof[(size_t)blockIdx.x * 16 + (k)] = clock64(); } while (0)

template <int IN_W>
__host__ __device__ constexpr size_t train_smem_bytes(int n_hidden) {
    return (size_t)IN_W * 128 + (size_t)(n_hidden - 1) * 8192 + 2048 + (size_t)IN_W * 256 + (size_t)n_hidden * 16384 + 2 * 16384 + 4096;
}
// TMEM columns: chain accumulator (64) + chain operand (32) + dW regions (IN_W, (H-1) x 64, 16)
__host__ __device__ constexpr uint32_t train_tmem_cols(int in_w, int n_hidden) { return 96u + (uint32_t)in_w + (uint32_t)(n_hidden - 1) * 64u + 16u; }

template <int N> struct TmemCols;
template <> struct TmemCols<32> {
    static __device__ __forceinline__ void ld(uint32_t t, uint32_t* r) { tc05::tmem_ld32(t, r); }
    static __device__ __forceinline__ void st_half(uint32_t t, const uint32_t* r) { tc05::tmem_st16(t, r); }
};
template <> struct TmemCols<16> {
    static __device__ __forceinline__ void ld(uint32_t t, uint32_t* r) { tc05::tmem_ld16(t, r); }
    static __device__ __forceinline__ void st_half(uint32_t t, const uint32_t* r) { tc05::tmem_st8(t, r); }
};

template <int IN_W, int TPR>
__global__ void __launch_bounds__(128 * TPR, 1) nrc_train_fused_kernel(const __grid_constant__ TrainArgs a) {
    using namespace tc05;
    constexpr int NT = 128 * TPR;
    constexpr int CPT = 64 / TPR;                 // accumulator columns of a hidden layer per thread
    constexpr int CH = CPT / 8;                   // 16-byte chunks (8 fp16 features) per thread and layer
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_mma, bar_dw[2], wbar;
    __shared__ uint32_t tmem_base_s, s_last;
    __shared__ float loss_red[4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, part = warp >> 2, r = quad * 32 + lane;
    const int H = a.n_hidden;
    uint8_t* w0_s = smem;
    uint8_t* wh_s = w0_s + IN_W * 128;
    uint8_t* wo_s = wh_s + (H - 1) * 8192;
    uint8_t* x_s = wo_s + 2048;
    uint8_t* act_s = x_s + IN_W * 256;
    uint8_t* dy_s = act_s + H * 16384;
    uint8_t* do_s = dy_s + 2 * 16384;

    NRC_PROF(0);
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (tid == 0) { mbar_init(&bar_mma, 1); mbar_init(&bar_dw[0], 1); mbar_init(&bar_dw[1], 1); mbar_init(&wbar, NT); fence_mbar_init(); }
    // everything above ran underneath the tail of the optimizer kernel launched before (programmatic dependent launch); the weights
    // it wrote are read from here on
    pdl_wait();
    pdl_trigger();          // the optimizer kernel that follows may be dispatched (its CTAs wait for this grid at their own pdl_wait)
    timeline_begin(a.tl, 0);
    // the weights travel to shared memory while the first tile is encoded (see nrc_forward2_kernel)
    copy_weights_kmajor_async(w0_s, a.params, kWidth, IN_W, tid, NT);
    for (int l = 1; l < H; l++) copy_weights_kmajor_async(wh_s + (l - 1) * 8192, a.params + IN_W * kWidth + (l - 1) * kWidth * kWidth, kWidth, kWidth, tid, NT);
    copy_weights_kmajor_async(wo_s, a.params + IN_W * kWidth + (H - 1) * kWidth * kWidth, kOutPad, kWidth, tid, NT);
    bool weights_pending = true;
    fence_before();
    __syncthreads();
    fence_after();

    const uint32_t tbase = tmem_base_s;
    const uint32_t tD = tbase, tA = tbase + 64, tW0 = tbase + 96, tWh = tW0 + IN_W, tWo = tWh + (uint32_t)(H - 1) * 64;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const uint32_t x_addr = smem_u32(x_s), w0_addr = smem_u32(w0_s), wh_addr = smem_u32(wh_s), wo_addr = smem_u32(wo_s);
    const uint32_t act_addr = smem_u32(act_s), dy_addr = smem_u32(dy_s), do_addr = smem_u32(do_s);
    const uint32_t idesc64 = make_idesc_f16(128, 64), idesc16 = make_idesc_f16(128, 16);                 // forward: B K-major
    const uint32_t idesc64b = make_idesc_f16(128, 64, 0, 1), idescxb = make_idesc_f16(128, IN_W, 0, 1);  // backward: B MN-major
    const uint32_t idw64 = make_idesc_f16(64, 64, 1, 1), idwx = make_idesc_f16(64, IN_W, 1, 1), idwo = make_idesc_f16(64, 16, 1, 1);
    const __half2* grid = reinterpret_cast<const __half2*>(a.params + a.n_mlp);
    const bool leader = tid == 0;
    const bool hashgrid = a.enc.pos_enc == POS_HASHGRID;
    const bool scatter = a.grid_grad != nullptr && hashgrid;
    const int L = a.enc.n_levels;
    const int lv_b = hashgrid ? (L * part) / TPR : 0, lv_e = hashgrid ? (L * (part + 1)) / TPR : L;
    uint32_t ph_mma = 0, ph_dw[2] = {0, 0};
    bool pend[2] = {false, false};
    const uint32_t n_tiles = a.n / kTile;
    const uint32_t row_off64 = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 16u;        // this record's row in a [128][64] tile
    uint8_t* my_x = x_s + (r >> 3) * (IN_W * 16) + (r & 7) * 16;
    uint32_t it = 0;
    NRC_PROF(1);

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
        const uint32_t row = tile * kTile + r;
        // the weight-gradient MMAs of the previous tile still read X and the activations
#pragma unroll
        for (int b = 0; b < 2; b++)
            if (pend[b]) { mbar_wait(&bar_dw[b], ph_dw[b]); ph_dw[b] ^= 1; pend[b] = false; }
        const float* rp = a.in + 5 * (size_t)row;
        const float x0 = rp[0], x1 = rp[1], x2 = rp[2], th = rp[3], ph = rp[4];
        float tg[3] = {0, 0, 0};
        if (part == 0) { const float* tp = a.target + 3 * (size_t)row; tg[0] = tp[0]; tg[1] = tp[1]; tg[2] = tp[2]; }
        {
            SmemRowPut put{my_x};
            // the optimizer will update exactly the entries this record touches: their 32-byte Adam records are prefetched into L2 with
            // the gathers (the DRAM round trips run underneath the MLP), so that nrc_grid_adam_kernel -- the other half of the step's
            // critical path, bound by its dependent gradient -> state loads -- finds them on chip
            const GridAdamState* pf = (NRC_TRAIN_PREFETCH && scatter) ? a.grid_state : nullptr;
            if (hashgrid || part == 0) encode_position<NRC_TRAIN_UNROLL>(a.enc, grid, x0, x1, x2, lv_b, lv_e, put, pf);
            if (part == TPR - 1) encode_direction_pad(a.enc, th, ph, put);
        }
        NRC_PROF(2);
        if (weights_pending) { cp_async_wait_all(); fence_proxy_async_smem(); mbar_arrive(&wbar); }
        fence_proxy_async_smem();
        fence_before();
        __syncthreads();
        NRC_PROF(3);
        if (leader) {
            if (weights_pending) mbar_wait(&wbar, 0);
            fence_after();
#pragma unroll
            for (int s = 0; s < IN_W / 16; s++)
                mma_f16_ss(tD, make_smem_desc(x_addr + s * 256, 128, IN_W * 16), make_smem_desc(w0_addr + s * 256, 128, IN_W * 16), idesc64, s > 0);
            mma_commit(&bar_mma);
        }
        weights_pending = false;
        mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
        fence_after();
        NRC_PROF(4);

        // ---------------- forward: hidden layers
        for (int l = 0; l < H; l++) {
            uint32_t acc[CPT], p[CPT / 2];
            TmemCols<CPT>::ld(tD + lane_base + part * CPT, acc);
            wait_ld();
            if (l == 0) {     // tcnn's ReLU is max(x, 0) in fp16: NaN (Q5) -> 0; cvt.relu would keep it
                const __half2 zero2 = __float2half2_rn(0.0f);
#pragma unroll
                for (int j = 0; j < CPT / 2; j++) {
                    uint32_t v = pack_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
                    __half2 m = __hmax2(*reinterpret_cast<__half2*>(&v), zero2);
                    p[j] = *reinterpret_cast<uint32_t*>(&m);
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT / 2; j++) p[j] = pack_relu_f16x2(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
            }
            TmemCols<CPT>::st_half(tA + lane_base + part * (CPT / 2), p);
            uint8_t* arow = act_s + l * 16384 + row_off64 + part * CH * 128;
#pragma unroll
            for (int c = 0; c < CH; c++) *reinterpret_cast<int4*>(arow + c * 128) = make_int4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
            wait_st();
            fence_proxy_async_smem();
            fence_before();
            __syncthreads();
            if (leader) {
                fence_after();
                if (l < H - 1) {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + l * 8192 + s * 256, 128, 1024), idesc64, s > 0);
                } else {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wo_addr + s * 256, 128, 1024), idesc16, s > 0);
                }
                mma_commit(&bar_mma);
            }
            mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
            fence_after();
        }

        NRC_PROF(5);
        // ---------------- output layer: loss and its gradient (relative_l2_luminance.h:40-88; n_total = batch * 3)
        if (part == 0) {
            uint32_t o[16];
            tmem_ld16(tD + lane_base, o);
            wait_ld();
            uint32_t ph16[8];
#pragma unroll
            for (int j = 0; j < 8; j++) ph16[j] = pack_f16x2(__uint_as_float(o[2 * j]), __uint_as_float(o[2 * j + 1]));
            float pr[3];
#pragma unroll
            for (int k = 0; k < 3; k++) pr[k] = __half2float(__float2half_rn(__uint_as_float(o[k])));
            const float n_total = (float)(a.n * 3u);
            const float lum = 0.299f * pr[0] + 0.587f * pr[1] + 0.114f * pr[2];
            const float denom = lum * lum + 0.01f;
            float loss = 0, gk[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float diff = pr[k] - tg[k];
                loss += diff * diff / denom / n_total;
                gk[k] = a.loss_scale * (2 * diff / denom) / n_total;
            }
            uint32_t g8[8] = {pack_f16x2(gk[0], gk[1]), pack_f16x2(gk[2], 0.0f), 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(tA + lane_base, g8);
            uint8_t* drow = do_s + (r >> 3) * 256 + (r & 7) * 16;
            *reinterpret_cast<int4*>(drow) = make_int4(g8[0], g8[1], 0, 0);
            *reinterpret_cast<int4*>(drow + 128) = make_int4(0, 0, 0, 0);
            if (a.out16) {
                int4* od = reinterpret_cast<int4*>(a.out16 + (size_t)row * kOutPad);
                od[0] = make_int4(ph16[0], ph16[1], ph16[2], ph16[3]); od[1] = make_int4(ph16[4], ph16[5], ph16[6], ph16[7]);
                int4* gd = reinterpret_cast<int4*>(a.dout16 + (size_t)row * kOutPad);
                gd[0] = make_int4(g8[0], g8[1], 0, 0); gd[1] = make_int4(0, 0, 0, 0);
            }
            // deterministic per-tile loss sum: warp shuffle tree, then the four quadrants in order
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, s);
            if (lane == 0) loss_red[quad] = loss;
            wait_st();
        }
        fence_proxy_async_smem();
        fence_before();
        __syncthreads();
        if (leader) {
            fence_after();
            a.loss_partials[tile] = ((loss_red[0] + loss_red[1]) + loss_red[2]) + loss_red[3];
            // dA_{H-1} = dOut (K = 16) x W_out read as [in][out]
            mma_f16_ts(tD, tA, make_smem_desc(wo_addr, 1024, 128), idesc64b, 0);
            mma_commit(&bar_mma);
            // dW_out^T [in][out] += A_{H-1}^T dOut   (M = 64 inputs, N = 16 outputs, K = 128 records)
#pragma unroll
            for (int s = 0; s < 8; s++)
                mma_f16_ss(tWo, make_smem_desc(act_addr + (H - 1) * 16384 + s * 2048, 1024, 128), make_smem_desc(do_addr + s * 512, 256, 128), idwo, (it > 0 || s > 0) ? 1u : 0u);
        }
        mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1;
        fence_after();

        NRC_PROF(6);
        // ---------------- backward: hidden layers
        const bool need_dx = scatter;
        for (int l = H - 1; l >= 0; l--) {
            const int b = l & 1;
            if (pend[b]) { mbar_wait(&bar_dw[b], ph_dw[b]); ph_dw[b] ^= 1; pend[b] = false; }      // dY buffer b is free again
            uint32_t acc[CPT], p[CPT / 2];
            int4 av[CH];
            const uint8_t* arow = act_s + l * 16384 + row_off64 + part * CH * 128;
#pragma unroll
            for (int c = 0; c < CH; c++) av[c] = *reinterpret_cast<const int4*>(arow + c * 128);
            TmemCols<CPT>::ld(tD + lane_base + part * CPT, acc);
            wait_ld();
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(av);
#pragma unroll
            for (int j = 0; j < CPT / 2; j++) {
                const float lo = (aw[j] & 0x0000ffffu) ? __uint_as_float(acc[2 * j]) : 0.0f;
                const float hi = (aw[j] & 0xffff0000u) ? __uint_as_float(acc[2 * j + 1]) : 0.0f;
                p[j] = pack_f16x2(lo, hi);
            }
            TmemCols<CPT>::st_half(tA + lane_base + part * (CPT / 2), p);
            uint8_t* drow = dy_s + b * 16384 + row_off64 + part * CH * 128;
#pragma unroll
            for (int c = 0; c < CH; c++) *reinterpret_cast<int4*>(drow + c * 128) = make_int4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
            wait_st();
            fence_proxy_async_smem();
            fence_before();
            __syncthreads();
            const bool chain = l > 0 || need_dx;
            if (leader) {
                fence_after();
                if (l > 0) {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(wh_addr + (l - 1) * 8192 + s * 2048, 1024, 128), idesc64b, s > 0);
                } else if (need_dx) {
#pragma unroll
                    for (int s = 0; s < 4; s++) mma_f16_ts(tD, tA + s * 8, make_smem_desc(w0_addr + s * (IN_W * 32), IN_W * 16, 128), idescxb, s > 0);
                }
                if (chain) mma_commit(&bar_mma);
                // dW_l [out][in] += dY_l^T A_{l-1}   (M = 64 outputs, N inputs, K = 128 records)
                const uint32_t acc0 = it > 0 ? 1u : 0u;
                if (l > 0) {
#pragma unroll
                    for (int s = 0; s < 8; s++)
                        mma_f16_ss(tWh + (uint32_t)(l - 1) * 64, make_smem_desc(dy_addr + b * 16384 + s * 2048, 1024, 128), make_smem_desc(act_addr + (l - 1) * 16384 + s * 2048, 1024, 128), idw64, (s > 0) ? 1u : acc0);
                } else {
#pragma unroll
                    for (int s = 0; s < 8; s++)
                        mma_f16_ss(tW0, make_smem_desc(dy_addr + b * 16384 + s * 2048, 1024, 128), make_smem_desc(x_addr + s * (IN_W * 32), IN_W * 16, 128), idwx, (s > 0) ? 1u : acc0);
                }
                mma_commit(&bar_dw[b]);
            }
            pend[b] = true;
            if (chain) { mbar_wait(&bar_mma, ph_mma); ph_mma ^= 1; fence_after(); }
        }

        NRC_PROF(7);
        // ---------------- hash-grid gradient: (half)weight * dL/dx, fp16x2 reductions (kernel_grid_backward, grid.h:215-320)
        if (need_dx) {
            __half2* gg = reinterpret_cast<__half2*>(a.grid_grad);
#pragma unroll 1
            for (int l = lv_b; l < lv_e; l++) {
                uint32_t t[2];
                tmem_ld2(tD + lane_base + 2 * l, t);
                wait_ld();
                const uint32_t gp = pack_f16x2(__uint_as_float(t[0]), __uint_as_float(t[1]));
                const __half2 g = *reinterpret_cast<const __half2*>(&gp);
                GridLevel c;
                grid_level_cell(a.enc, l, x0, x1, x2, c);
                __half2* base = gg + a.enc.level_offset[l];
#pragma unroll
                for (int k = 0; k < 8; k++) red_add_f16x2(base + c.idx[k], __hmul2(__half2half2(__float2half_rn(c.w[k])), g));
            }
            // (every thread's reads of tD are complete -- wait::ld -- before it reaches the next tile's first barrier)
        }
    }
    NRC_PROF(8);
    if (weights_pending) { cp_async_wait_all(); mbar_arrive(&wbar); }
#pragma unroll
    for (int b = 0; b < 2; b++)
        if (pend[b]) { mbar_wait(&bar_dw[b], ph_dw[b]); ph_dw[b] ^= 1; pend[b] = false; }
    fence_after();

    // ---------------- weight-gradient partial of this CTA.  M = 64 accumulators: row 16 * q + i (i < 16) of D lives in TMEM lane 32 * q + i
    if (it > 0) {
        float* out = a.dw_partials + (size_t)blockIdx.x * a.n_mlp;
        const int drow = quad * 16 + lane;
        for (int m = 0; m <= H; m++) {
            const int bw = m == 0 ? IN_W : m < H ? kWidth : kOutPad;
            const uint32_t tW = m == 0 ? tW0 : m < H ? tWh + (uint32_t)(m - 1) * 64 : tWo;
            const size_t moff = m == 0 ? 0 : (size_t)IN_W * kWidth + (size_t)(m - 1) * kWidth * kWidth;
            for (int c = part * 16; c < bw; c += 16 * TPR) {
                uint32_t v[16];
                tmem_ld16(tW + lane_base + c, v);
                wait_ld();
                if (lane < 16) {
                    if (m < H) {
                        float4* dst = reinterpret_cast<float4*>(out + moff + (size_t)drow * bw + c);
#pragma unroll
                        for (int q = 0; q < 4; q++) dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    } else {      // output layer accumulated transposed: D[in][out] -> W_out[out][in]
#pragma unroll
                        for (int q = 0; q < 16; q++) out[moff + (size_t)(c + q) * kWidth + drow] = __uint_as_float(v[q]);
                    }
                }
            }
        }
    }
    NRC_PROF(9);
    // ---------------- Trainer::loss (trainer.h:205-207): the CTA that finishes last adds the tile partials in tile order
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.done_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        float s = 0;
        for (uint32_t i0 = 0; i0 < n_tiles; i0 += 32) {
            const float v = (i0 + lane < n_tiles) ? __ldcg(a.loss_partials + i0 + lane) : 0.0f;
#pragma unroll
            for (int j = 0; j < 32; j++) { const float vj = __shfl_sync(0xffffffffu, v, j); if (i0 + j < n_tiles) s += vj; }
        }
        if (lane == 0) { *a.loss_out = s; *a.done_counter = 0u; }
    }
    fence_before();
    __syncthreads();
    NRC_PROF(10);
    timeline_end(a.tl, 0);
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

// fp32 sum of the per-CTA weight-gradient partials, fixed order (deterministic): a CTA owns 64 parameters, four thread groups
// add a quarter of the chunks each, the four partial sums are added in group order
__global__ void __launch_bounds__(256) nrc_reduce_partials_kernel(const float* __restrict__ partials, uint32_t n_chunks, uint32_t n_mlp, float* __restrict__ out,
                                                                  unsigned long long* tl = nullptr) {
    __shared__ float s[4][64];
    timeline_begin(tl, 0);
    const uint32_t pl = threadIdx.x & 63, slice = threadIdx.x >> 6, i = blockIdx.x * 64 + pl;
    const uint32_t per = (n_chunks + 3) / 4, c0 = slice * per, c1 = min(n_chunks, c0 + per);
    float g = 0;
    if (i < n_mlp) {
#pragma unroll 8
        for (uint32_t c = c0; c < c1; c++) g += partials[(size_t)c * n_mlp + i];
    }
    s[slice][pl] = g;
    __syncthreads();
    if (slice == 0 && i < n_mlp) out[i] = ((s[0][pl] + s[1][pl]) + s[2][pl]) + s[3][pl];
    timeline_end(tl, 0);
}

// materialise the fp16 MLP gradient (what tcnn keeps in Trainer::param_gradients) without running the optimizer
__global__ void __launch_bounds__(256) nrc_partials_to_half_kernel(const float* __restrict__ partials, uint32_t n_chunks, uint32_t n_mlp, __half* __restrict__ out) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_mlp) return;
    float g = 0;
    for (uint32_t c = 0; c < n_chunks; c++) g += partials[(size_t)c * n_mlp + i];
    out[i] = __float2half_rn(g);
}

// ---------------------------------------------------------------------------------------------- optimizer
struct OptArgs {
    uint64_t n_params, n_mlp;
    float* master;              // network weights only (SoA, n_mlp)
    float* m1;
    float* m2;
    uint32_t* steps;
    GridAdamState* grid_state;  // one 32-byte record per hash-grid entry
    __half* w16;
    __half* ema16;
    __half* grad16;
    const float* partials;
    uint32_t n_chunks;
    uint32_t mlp_blocks;        // nrc_adam_kernel: CTAs [0, mlp_blocks) own 64 network weights each
    uint64_t grid_begin, grid_end;   // parameter range the hash-grid CTAs cover (the whole encoding; a data-parallel rank's own slice)
    float lr, beta1, beta2, eps, l2_reg, loss_scale, ema_decay, ema_debias_old, ema_debias_new, log2_beta1, log2_beta2;
    unsigned long long* tl;     // optional device-side timeline slots (development aid)
    unsigned long long* tl_ema;
};

// adam.h:48-121 for one parameter whose state is already in registers; returns the new fp32 master weight
__device__ __forceinline__ float adam_update(const OptArgs& a, float wfp, float& m1, float& m2, uint32_t& st, float gradient) {
    const float gsq = gradient * gradient;
    m1 = a.beta1 * m1 + (1 - a.beta1) * gradient;
    m2 = a.beta2 * m2 + (1 - a.beta2) * gsq;
    st++;
    // beta^t as exp2(t * log2 beta): MUFU.EX2 instead of powf's ~100-instruction slow path (relative error ~1e-7)
    const float lr = a.lr * (sqrtf(1 - exp2f((float)st * a.log2_beta2)) / (1 - exp2f((float)st * a.log2_beta1)));
    const float eff = fminf(fmaxf(lr / (sqrtf(m2) + a.eps), 0.0f), 3.402823466e+38f);
    return wfp - eff * m1;
}

// The optimizer runs as two kernels (adam.h:48-121, ema.h:63-76):
//   nrc_adam_kernel      ONE launch on the training stream: the first n_mlp / 64 CTAs own 64 network weights each (fixed-order sum of
//                        the per-CTA weight-gradient partials, dense Adam, EMA); the remaining CTAs run Adam on the hash-grid entries
//                        whose gradient is non-zero (adam.h:77-80) -- the only part of the optimizer the NEXT training step depends on
//                        (it reads the fp16 working weights).  (Round 2 measured the network part as its own kernels on a side
//                        stream: they were only dispatched once every CTA of the encoding grid had been, 25 us on the critical path.)
//   nrc_grid_ema_kernel  hash-grid entries: dense EMA of the fp16 weights (what Inference() reads); nothing in a training step reads
//                        it, so it runs on a side stream underneath the next step's forward / backward kernel: a persistent grid of one
//                        small CTA per SM that fits next to that kernel's CTA.
// The touched hash-grid entries are few (a quarter at the reference's batch size) and scattered, so each warp first compacts them
// (ballot + popc ranks into a shared list) and then runs Adam densely, one touched ENTRY (two features, one 32-byte state sector) per
// lane, instead of executing mostly-predicated-off copies of the update.  The new fp16 weights are written entry by entry; nothing
// else of the fp16 weight vector is read.  Consumed gradients are re-zeroed here, so the next step needs no 28.5 MB memset.
#ifndef NRC_OPT_MIN_BLOCKS
#define NRC_OPT_MIN_BLOCKS 6
#endif
#ifndef NRC_OPT_STREAM_STATE
#define NRC_OPT_STREAM_STATE 0
#endif
__global__ void __launch_bounds__(256, NRC_OPT_MIN_BLOCKS) nrc_adam_kernel(const __grid_constant__ OptArgs a) {
    __shared__ uint32_t s_g[8][128];       // per warp: compacted gradients (half2 bits) of the touched entries ... (network CTAs: partial sums)
    __shared__ uint16_t s_el[8][128];      // ... and their entry index inside the warp's 128-entry span
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_scale = 1.0f / a.loss_scale;
    pdl_wait();             // the gradients of the training kernel launched before (programmatic dependent launch)
    pdl_trigger();          // the next training kernel's CTAs may be dispatched as this grid drains; they wait for its completion themselves
    timeline_begin(a.tl, 0);
    if (blockIdx.x < a.mlp_blocks) {
        // ---- network weights: 64 per CTA; four thread groups add a quarter of the partials each, in chunk order (deterministic)
        float* s = reinterpret_cast<float*>(&s_g[0][0]);          // [4][64]
        const uint32_t pl = threadIdx.x & 63, slice = threadIdx.x >> 6, i = blockIdx.x * 64 + pl;
        const uint32_t per = (a.n_chunks + 3) / 4, c0 = slice * per, c1 = min(a.n_chunks, c0 + per);
        float gsum = 0;
#pragma unroll 8
        for (uint32_t c = c0; c < c1; c++) gsum += a.partials[(size_t)c * a.n_mlp + i];
        s[slice * 64 + pl] = gsum;
        __syncthreads();
        if (slice == 0) {
            const float acc = ((s[pl] + s[64 + pl]) + s[128 + pl]) + s[192 + pl];
            const __half g16 = __float2half_rn(acc);                       // tcnn keeps gradients in fp16 (trainer.h:322-336)
            float mw = a.master[i], m1 = a.m1[i], m2 = a.m2[i];
            uint32_t st = a.steps[i];
            const float gradient = __half2float(g16) * inv_scale + a.l2_reg * mw;
            mw = adam_update(a, mw, m1, m2, st, gradient);
            const __half w = __float2half_rn(mw);
            a.master[i] = mw; a.m1[i] = m1; a.m2[i] = m2; a.steps[i] = st;
            a.grad16[i] = g16; a.w16[i] = w;
            const float filtered = (__half2float(a.ema16[i]) * a.ema_decay * a.ema_debias_old + __half2float(w) * (1 - a.ema_decay)) * a.ema_debias_new;
            a.ema16[i] = __float2half_rn(filtered);
        }
        return;
    }
    const uint32_t gb = blockIdx.x - a.mlp_blocks;
    const uint64_t i0 = a.grid_begin + ((uint64_t)gb * 256 + threadIdx.x) * 8;
    const uint64_t warp_i0 = a.grid_begin + ((uint64_t)gb * 256 + warp * 32) * 8;
    if (warp_i0 >= a.grid_end) return;     // whole warp out of range
    union V8 { int4 v; uint32_t u[4]; };
    V8 g;
    g.v = make_int4(0, 0, 0, 0);
    if (i0 < a.grid_end) g.v = *reinterpret_cast<const int4*>(a.grad16 + i0);
    uint32_t base = 0, my_rank[4];
    bool mine[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        mine[j] = (g.u[j] & 0x7fff7fffu) != 0u;              // at least one of the entry's two gradients is non-zero (+-0 both skip)
        const uint32_t b = __ballot_sync(0xffffffffu, mine[j]);
        my_rank[j] = base + __popc(b & ((1u << lane) - 1));
        base += __popc(b);
    }
    if (base) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (mine[j]) { s_el[warp][my_rank[j]] = (uint16_t)(lane * 4 + j); s_g[warp][my_rank[j]] = g.u[j]; }
        if (mine[0] | mine[1] | mine[2] | mine[3]) *reinterpret_cast<int4*>(a.grad16 + i0) = make_int4(0, 0, 0, 0);   // consumed
        __syncwarp();
        GridAdamState* st0 = a.grid_state + ((warp_i0 - a.n_mlp) >> 1);          // first entry of the warp's span
        uint32_t* w0 = reinterpret_cast<uint32_t*>(a.w16 + warp_i0);
        for (uint32_t r = lane; r < base; r += 32) {
            const uint32_t el = s_el[warp][r];
            float4* sp = reinterpret_cast<float4*>(st0 + el);
#if NRC_OPT_STREAM_STATE
            float4 s0 = __ldcs(sp);                             // master.xy, m1.xy   (evict-first: the state streams through L2 once per step)
            float4 s1 = __ldcs(sp + 1);                         // m2.xy, steps.xy
#else
            float4 s0 = sp[0];                                  // master.xy, m1.xy
            float4 s1 = sp[1];                                  // m2.xy, steps.xy
#endif
            const uint32_t gbits = s_g[warp][r];
            const __half2 gh = *reinterpret_cast<const __half2*>(&gbits);
            const float g0 = __low2float(gh), g1 = __high2float(gh);
            uint32_t t0 = __float_as_uint(s1.z), t1 = __float_as_uint(s1.w);
            if (g0 != 0.0f) s0.x = adam_update(a, s0.x, s0.z, s1.x, t0, g0 * inv_scale);
            if (g1 != 0.0f) s0.y = adam_update(a, s0.y, s0.w, s1.y, t1, g1 * inv_scale);
            s1.z = __uint_as_float(t0); s1.w = __uint_as_float(t1);
#if NRC_OPT_STREAM_STATE
            __stcs(sp, s0); __stcs(sp + 1, s1);
#else
            sp[0] = s0; sp[1] = s1;
#endif
            // fp16 working weights of the entry; a feature whose own gradient is zero keeps its master weight, hence its fp16 bits
            w0[el] = tc05::pack_f16x2(s0.x, s0.y);
        }
    }
    timeline_end(a.tl, 0);
}

// dense EMA of the hash-grid part of the fp16 weight vector (ema.h:63-76): persistent grid-stride loop, four 16-byte vectors of each
// array in flight per thread
__global__ void __launch_bounds__(256) nrc_grid_ema_kernel(const __grid_constant__ OptArgs a) {
    timeline_begin(a.tl_ema, 0);
    grid_ema_pass(a.w16 + a.n_mlp, a.ema16 + a.n_mlp, (a.n_params - a.n_mlp) / 8, a.ema_decay, a.ema_debias_old, a.ema_debias_new,
                  (uint64_t)blockIdx.x * 256 + threadIdx.x, (uint64_t)gridDim.x * 256);
    timeline_end(a.tl_ema, 0);
}

// (un)packing between the tcnn parameter order and the per-entry Adam records (nrc_get_params / nrc_set_params_fp32)
__global__ void __launch_bounds__(256) nrc_grid_state_scatter_kernel(GridAdamState* st, uint64_t n_entries, int field, const float* __restrict__ src) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_entries) return;
    float* f = reinterpret_cast<float*>(st + i) + 2 * field;
    f[0] = src[2 * i]; f[1] = src[2 * i + 1];
}
__global__ void __launch_bounds__(256) nrc_grid_state_gather_kernel(const GridAdamState* st, uint64_t n_entries, int field, float* __restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_entries) return;
    const float* f = reinterpret_cast<const float*>(st + i) + 2 * field;
    if (field == 3) { dst[2 * i] = (float)__float_as_uint(f[0]); dst[2 * i + 1] = (float)__float_as_uint(f[1]); }
    else { dst[2 * i] = f[0]; dst[2 * i + 1] = f[1]; }
}

// ---------------------------------------------------------------------------------------------- data-parallel training over peer memory
// Replaces the NCCL all-reduce a data-parallel tcnn would need with a SHARDED optimizer step over NVLink peer memory
// (buffers shared with cudaIpc).  Every rank owns 1/world of the hash-grid entries:
//   1. nrc_peer_gather_kernel   reduce-scatter: the rank reads ITS slice of every peer's fp16 gradient straight out of the peer's HBM
//                               (P2P loads), sums in rank order in fp32, keeps the sum in its own gradient buffer and writes zeros back
//                               over the peers' words it consumed (so nobody needs a 28.5 MB memset).  The small fp32 MLP gradient is
//                               summed redundantly by every rank, in rank order.
//   2. nrc_adam_kernel          Adam on the rank's own slice only (plus the network weights, redundantly and identically everywhere):
//                               at eight ranks the union of the touched entries is ~90 % of the tables, a replicated optimizer would
//                               stream 4x the single-GPU state per step; sharded, every rank updates 1/8 of it.
//   3. nrc_peer_publish_kernel  all-gather: the rank stores its updated fp16 weight slice into every peer's weight vector (P2P stores).
// Replicas therefore hold bit-identical fp16 weights (working and EMA) after every step; the fp32 master weights and Adam moments of
// a slice live on its owner only.  Synchronisation: three flag rounds in peer-visible memory (system-scope release / acquire):
// A "my backward pass is complete, my previous EMA pass has read the weights", B "I am done reading and clearing your gradients",
// C "my weight slice has landed in your buffer".
constexpr int kMaxPeers = 8;
struct PeerArgs {
    int rank, world;
    int4* grad[kMaxPeers];             // hash-grid gradient of every rank (own buffer at [rank]), 8 halfs per int4
    int4* w16[kMaxPeers];              // hash-grid fp16 weights of every rank
    const float* mlp[kMaxPeers];       // fp32 MLP gradient of every rank
    uint32_t* flags[kMaxPeers];        // flags[q]: rank q's flag words [3][kMaxPeers]; slot [round][p] is written by rank p
    float* mlp_sum;                    // local: sum over the ranks
    uint64_t n_vec;                    // int4 words of the hash-grid gradient
    uint64_t slice_begin, slice_end;   // this rank's slice, in int4 words
    uint32_t n_mlp;
    uint32_t token;                    // exchange counter, identical on every rank
    unsigned int* done_counter;        // local
    unsigned long long* tl;            // optional device-side timeline slot (development aid)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

template <int WORLD>      // upper bound of a.world: 2, 4 or 8 (sizes the register tile of in-flight peer loads)
__global__ void __maxnreg__(48) nrc_peer_gather_kernel(const __grid_constant__ PeerArgs a) {
    timeline_begin(a.tl, 0);
    // ---- round A: every rank's backward pass has finished (stream order makes that true for this rank at kernel start)
    if (blockIdx.x == 0 && threadIdx.x < a.world) st_release_sys(a.flags[threadIdx.x] + a.rank, a.token);
    if (threadIdx.x < a.world) while ((int32_t)(ld_acquire_sys(a.flags[a.rank] + threadIdx.x) - a.token) < 0) { }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = gtid; i < a.n_mlp; i += stride) {
        float sum = 0;
        for (int p = 0; p < a.world; p++) sum += a.mlp[p][i];
        a.mlp_sum[i] = sum;
    }
    // 8 / WORLD independent 16-byte words per peer, thread and round: all peer loads of a round are in flight together (a remote load
    // costs ~1 us of NVLink latency); the register tile stays at 32 registers
    constexpr int kU = 8 / WORLD;
    const int4 zero = make_int4(0, 0, 0, 0);
    for (uint64_t i0 = a.slice_begin + gtid; i0 < a.slice_end; i0 += stride * kU) {
        int4 v[WORLD][kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const uint64_t i = i0 + (uint64_t)u * stride;
#pragma unroll
            for (int p = 0; p < WORLD; p++)
                if (p < a.world) v[p][u] = i < a.slice_end ? a.grad[p][i] : zero;
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t any = 0;
#pragma unroll
            for (int p = 0; p < WORLD; p++) {
                if (p >= a.world) break;
                const uint32_t nz = (uint32_t)(v[p][u].x | v[p][u].y | v[p][u].z | v[p][u].w) & 0x7fff7fffu;
                any |= nz;
                const __half2* h = reinterpret_cast<const __half2*>(&v[p][u]);
#pragma unroll
                for (int k = 0; k < 4; k++) { const float2 f = __half22float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
                if (nz && p != a.rank && i < a.slice_end) a.grad[p][i] = zero;          // consumed: the peer's next scatter starts from zero
            }
            if (i >= a.slice_end || any == 0) continue;                                 // untouched on every rank (+-0): nothing to do
            int4 r;
            __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
            a.grad[a.rank][i] = r;                                                      // the sum stays with the owner of the slice
        }
    }
    // ---- round B: the last block of this rank tells every rank that this rank no longer touches their gradient buffers
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(a.done_counter, 1u) == gridDim.x - 1) {
            *a.done_counter = 0;
            __threadfence_system();
            for (int q = 0; q < a.world; q++) st_release_sys(a.flags[q] + kMaxPeers + a.rank, a.token);
        }
    }
    timeline_end(a.tl, 0);
}

// all-gather of the updated fp16 weights: the rank's slice goes to every peer.  Starts by waiting for round B (every peer has finished
// reading this rank's gradients, i.e. the next scatter may write them) and ends with round C.
__global__ void __maxnreg__(48) nrc_peer_publish_kernel(const __grid_constant__ PeerArgs a) {
    timeline_begin(a.tl, 0);
    if (threadIdx.x < a.world) while ((int32_t)(ld_acquire_sys(a.flags[a.rank] + kMaxPeers + threadIdx.x) - a.token) < 0) { }
    __syncthreads();
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = a.slice_begin + gtid; i0 < a.slice_end; i0 += stride * 4) {
        int4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const uint64_t i = i0 + (uint64_t)u * stride; if (i < a.slice_end) v[u] = a.w16[a.rank][i]; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            if (i >= a.slice_end) break;
            for (int p = 0; p < a.world; p++)
                if (p != a.rank) a.w16[p][i] = v[u];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(a.done_counter, 1u) == gridDim.x - 1) {
            *a.done_counter = 0;
            __threadfence_system();
            for (int q = 0; q < a.world; q++) st_release_sys(a.flags[q] + 2 * kMaxPeers + a.rank, a.token);
        }
    }
    timeline_end(a.tl, 0);
}

// ONE kernel for steps 1-3 (default; the three separate kernels above remain as the step-by-step form the tests compare it with, bit for
// bit): a warp owns a span of 32 16-byte words (128 hash-grid entries) of the rank's slice.  It receives the span from every peer's
// gradient, sums in rank order, clears the consumed words at their owners, compacts
// the touched entries (ballot + popc), runs Adam on their 32-byte state records (local HBM) and stores the updated fp16 weight words
// into its own and every peer's weight vector -- reduce-scatter, optimizer and all-gather pipelined span by span instead of three
// bulk-synchronous phases (2 ranks: 51 + 36 + 28 us and two launch gaps -> one launch).  The first CTAs update the network weights from
// the sum of the ranks' fp32 MLP gradients, identically on every rank.  Flag rounds: A as above; B and C collapse into one final round.
// The peer loads are the latency that bounds the kernel (NVLink round trip ~2 us idle, ~10 us loaded; ~300 GB/s arrive per direction on
// this box at these message sizes, NCCL's send / recv does no better), and registers bound how many loads a thread can keep in flight.
// So the gradient words travel with cp.async (LDGSTS: peer HBM -> this SM's shared memory, no register staging): a persistent CTA
// works through blocks of 256 words and keeps DEPTH - 1 blocks of every peer in flight; each thread only ever reads the slots it
// filled itself, so cp.async.wait_group is the only synchronisation of the pipeline.
template <int WORLD> __host__ __device__ constexpr int peer_adam_depth() { return WORLD <= 2 ? 4 : WORLD <= 4 ? 3 : 2; }
template <int WORLD> __host__ __device__ constexpr size_t peer_adam_smem_bytes() { return (size_t)peer_adam_depth<WORLD>() * WORLD * 256 * sizeof(int4); }

template <int WORLD>
__global__ void __launch_bounds__(256) nrc_peer_adam_kernel(const __grid_constant__ OptArgs a, const __grid_constant__ PeerArgs pa) {
    constexpr int DEPTH = peer_adam_depth<WORLD>();
    extern __shared__ __align__(16) int4 stage[];          // [DEPTH][WORLD][256]
    __shared__ uint32_t s_g[8][128];       // per warp: compacted summed gradients (half2 bits) of the touched entries ...
    __shared__ uint16_t s_el[8][128];      // ... their entry index inside the warp's 128-entry span ...
    __shared__ uint32_t s_w[8][128];       // ... and the new fp16 weights by rank
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float inv_scale = 1.0f / a.loss_scale;
    timeline_begin(a.tl, 0);
    // ---- round A: every rank's backward pass (and its previous EMA pass over the weights) has finished
    if (blockIdx.x == 0 && threadIdx.x < pa.world) st_release_sys(pa.flags[threadIdx.x] + pa.rank, pa.token);
    if (threadIdx.x < pa.world) while ((int32_t)(ld_acquire_sys(pa.flags[pa.rank] + threadIdx.x) - pa.token) < 0) { }
    __syncthreads();
    if (blockIdx.x < a.mlp_blocks) {
        // ---- network weights: 64 per CTA, gradient = sum of the ranks' fp32 gradients in rank order (identical on every rank)
        if (threadIdx.x < 64) {
            const uint32_t i = blockIdx.x * 64 + threadIdx.x;
            float acc = 0;
            for (int p = 0; p < pa.world; p++) acc += pa.mlp[p][i];
            const __half g16 = __float2half_rn(acc);                       // tcnn keeps gradients in fp16 (trainer.h:322-336)
            float mw = a.master[i], m1 = a.m1[i], m2 = a.m2[i];
            uint32_t st = a.steps[i];
            const float gradient = __half2float(g16) * inv_scale + a.l2_reg * mw;
            mw = adam_update(a, mw, m1, m2, st, gradient);
            const __half w = __float2half_rn(mw);
            a.master[i] = mw; a.m1[i] = m1; a.m2[i] = m2; a.steps[i] = st;
            a.grad16[i] = g16; a.w16[i] = w;
            const float filtered = (__half2float(a.ema16[i]) * a.ema_decay * a.ema_debias_old + __half2float(w) * (1 - a.ema_decay)) * a.ema_debias_new;
            a.ema16[i] = __float2half_rn(filtered);
        }
    } else {
        const uint64_t gb = blockIdx.x - a.mlp_blocks, n_ctas = gridDim.x - a.mlp_blocks;
        const uint64_t n_blocks = (pa.slice_end - pa.slice_begin + 255) / 256;
        union V8 { int4 v; uint32_t u[4]; };
        const int4 zero = make_int4(0, 0, 0, 0);
        auto issue = [&](uint64_t blk, int buf) {           // this thread's word of block `blk`, from every rank's gradient -> its own slots
            if (blk < n_blocks) {
                const uint64_t wd = min(pa.slice_begin + blk * 256 + threadIdx.x, pa.slice_end - 1);
#pragma unroll
                for (int p = 0; p < WORLD; p++)
                    if (p < pa.world) tc05::cp_async16(&stage[(buf * WORLD + p) * 256 + threadIdx.x], &pa.grad[p][wd]);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int d = 0; d < DEPTH - 1; d++) issue(gb + d * n_ctas, d);
        uint32_t it = 0;
        for (uint64_t blk = gb; blk < n_blocks; blk += n_ctas, it++) {
            issue(blk + (DEPTH - 1) * n_ctas, (it + DEPTH - 1) % DEPTH);
            asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
            const int buf = it % DEPTH;
            const uint64_t warp_word = pa.slice_begin + blk * 256 + warp * 32, word = warp_word + lane;
            if (warp_word >= pa.slice_end) continue;                          // whole span out of range (warp-uniform)
            const bool valid = word < pa.slice_end;
            V8 w;
            w.v = valid ? pa.w16[pa.rank][word] : zero;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t any = 0;
#pragma unroll
            for (int p = 0; p < WORLD; p++) {
                if (p >= pa.world) break;
                const int4 v = valid ? stage[(buf * WORLD + p) * 256 + threadIdx.x] : zero;
                const uint32_t nz = (uint32_t)(v.x | v.y | v.z | v.w) & 0x7fff7fffu;
                any |= nz;
                const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int k = 0; k < 4; k++) { const float2 f = __half22float2(h[k]); acc[2 * k] += f.x; acc[2 * k + 1] += f.y; }
                if (nz) pa.grad[p][word] = zero;                            // consumed (own and peers'): the next scatter starts from zero
            }
            V8 g;
            __half2* gh = reinterpret_cast<__half2*>(&g.v);
#pragma unroll
            for (int k = 0; k < 4; k++) gh[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
            if (!any) g.v = zero;
            uint32_t base = 0, my_rank[4];
            bool mine[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                mine[j] = (g.u[j] & 0x7fff7fffu) != 0u;
                const uint32_t b = __ballot_sync(0xffffffffu, mine[j]);
                my_rank[j] = base + __popc(b & ((1u << lane) - 1));
                base += __popc(b);
            }
            if (!base) continue;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (mine[j]) { s_el[warp][my_rank[j]] = (uint16_t)(lane * 4 + j); s_g[warp][my_rank[j]] = g.u[j]; }
            __syncwarp();
            GridAdamState* st0 = a.grid_state + warp_word * 4;               // first entry of the span
            for (uint32_t r = lane; r < base; r += 32) {
                const uint32_t el = s_el[warp][r];
                float4* sp = reinterpret_cast<float4*>(st0 + el);
                float4 s0 = sp[0];                                  // master.xy, m1.xy
                float4 s1 = sp[1];                                  // m2.xy, steps.xy
                const uint32_t gbits = s_g[warp][r];
                const __half2 g2 = *reinterpret_cast<const __half2*>(&gbits);
                const float g0 = __low2float(g2), g1 = __high2float(g2);
                uint32_t t0 = __float_as_uint(s1.z), t1 = __float_as_uint(s1.w);
                if (g0 != 0.0f) s0.x = adam_update(a, s0.x, s0.z, s1.x, t0, g0 * inv_scale);
                if (g1 != 0.0f) s0.y = adam_update(a, s0.y, s0.w, s1.y, t1, g1 * inv_scale);
                s1.z = __uint_as_float(t0); s1.w = __uint_as_float(t1);
                sp[0] = s0; sp[1] = s1;
                s_w[warp][r] = tc05::pack_f16x2(s0.x, s0.y);
            }
            __syncwarp();
            if (mine[0] | mine[1] | mine[2] | mine[3]) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (mine[j]) w.u[j] = s_w[warp][my_rank[j]];
                for (int p = 0; p < pa.world; p++) pa.w16[p][word] = w.v;      // all-gather of the updated word (own copy included)
            }
            __syncwarp();                                                      // the next block re-uses the warp's shared lists
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    // ---- final round: the last block of this rank tells every rank that its loads, clears and weight stores are globally visible
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(pa.done_counter, 1u) == gridDim.x - 1) {
            *pa.done_counter = 0;
            __threadfence_system();
            for (int q = 0; q < pa.world; q++) {
                st_release_sys(pa.flags[q] + kMaxPeers + pa.rank, pa.token);
                st_release_sys(pa.flags[q] + 2 * kMaxPeers + pa.rank, pa.token);
            }
        }
    }
    timeline_end(a.tl, 0);
}

// the next forward pass may gather the weights once every rank's slice has landed (round C)
__global__ void nrc_peer_wait_kernel(const uint32_t* flags, int world, uint32_t token) {
    if ((int)threadIdx.x < world) while ((int32_t)(ld_acquire_sys(flags + 2 * kMaxPeers + threadIdx.x) - token) < 0) { }
}

}  // namespace nrchpm
