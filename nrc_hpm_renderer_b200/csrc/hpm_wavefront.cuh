// Path-regeneration form of the gen_rays pass (gen_rays.comp + prep_infer_rays.comp): the same per-pixel arithmetic and the same
// per-pixel RNG stream as hpm_gen_rays_kernel -- every output is bit-identical -- but a lane whose path has ended does not idle until
// the longest path of its warp has ended; it takes the next path from a queue ("warp-level primitives for compaction of terminated
// paths"):
//
//   hpm_wf_primary_kernel   one pixel per thread (coherent): camera ray, analytic sky test, FindEntryExit.  Sky pixels are finished here;
//                           the pixels that reach the volume get a compact path record (ballot + popc, one atomic per warp)
//   hpm_wf_paths_kernel     persistent warps over the path queue.  One loop iteration = one bounce of TracePath (DeltaTrack, TraceScene,
//                           NewRayDir, Russian roulette) for every lane that holds a path; lanes without one are refilled from the
//                           queue at the bounce boundary (chunks of 32 queue entries per warp-level atomic, ranks by ballot + popc).
//                           When the queue has run dry and fewer than `spill_below` lanes of a warp are still alive, the warp writes
//                           the survivors back as path records of a SECOND queue and retires; the next launch of the same kernel
//                           regroups them into full warps.  The last launch runs every path to its end.
//
// What this buys: in the long-path configuration (BASELINE config 4: primaryRayLength 4, primaryRayProb .75) the pixel-per-thread kernel
// keeps a warp alive until its longest path has ended.  What it does not change: the lanes lost INSIDE the tracking loops of one bounce
// (a warp runs the delta-tracking loop until its last lane has found a real collision).  A finer-grained wavefront (one queue per phase
// of a bounce, lanes refilled inside the tracking loops) was built and measured in round 2: bit-identical, and SLOWER on B200 at both
// configurations (0.74 vs 0.41 ms, 6.3 vs 4.1 ms) -- a 1080p frame has ~5e5 paths for 3e5 resident lanes, so every phase launch is a
// queue that runs dry after 2-5 entries per lane, and the refill code runs on nearly every loop iteration with one or two lanes active.
#pragma once
#include "hpm_kernels.cuh"

namespace nrchpm {

struct WfState {            // path records, SoA over compact path ids
    uint32_t* pixel;        // x | y << 16
    float* rng;
    float* cur;             // [3][n]
    float* dir;             // [3][n]
    float* light;           // [3][n]
    float* factor;
    uint32_t* bounce;       // loop index i of TracePath | did_scatter << 16
    uint32_t n;             // stride
};

// Device counters of one frame (uint32, zeroed by ONE memset per frame): round r owns words 2r (queue length) and 2r + 1 (queue head)
constexpr int kWfMaxRounds = 6;
constexpr int kWfRounds = 3;                 // launches of the path kernel per frame (measured, scripts/tune_wavefront.py)
constexpr uint32_t kWfSpillBelow = 16;       // a warp of an early round retires when the queue is dry and fewer lanes than this are alive

struct WfArgs {
    SceneDev sc; CameraDev cam; RenderCfgDev cfg;
    float4 frame_random;
    float4* primary_color; float* info; float* origin; float* dir; float* infer_in;
    uint32_t* infer_filter; uint32_t* active_list; uint32_t* active_count;
    unsigned long long* lookups;
    WfState st;
};

namespace hpmdev {

__device__ __forceinline__ V3 ld3(const float* a, uint32_t n, uint32_t i) { return mk(a[i], a[n + i], a[2 * n + i]); }
__device__ __forceinline__ void st3(float* a, uint32_t n, uint32_t i, V3 v) { a[i] = v.x; a[n + i] = v.y; a[2 * n + i] = v.z; }

// append `item` for the lanes with `flag` set to a queue: ballot + popc ranks, one atomic per warp (all 32 lanes must call)
__device__ __forceinline__ void warp_push(bool flag, uint32_t item, uint32_t* items, uint32_t* count) {
    const uint32_t lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31;
    const uint32_t ballot = __ballot_sync(0xffffffffu, flag);
    if (!ballot) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (flag) items[base + __popc(ballot & ((1u << lane) - 1))] = item;
}

// end of a pixel's path (gen_rays.comp:44-51, 82-99 + prep_infer_rays.comp:26-46): images, query record, filter flag
__device__ __forceinline__ void wf_finish(const WfArgs& a, const Tracker& c, uint32_t x, uint32_t y, V3 cur, V3 dir, V3 light, float factor, bool did_scatter) {
    const uint32_t W = a.cfg.width, H = a.cfg.height;
    const size_t p = (size_t)y * W + x;
    if (a.origin) { a.origin[3 * p + 0] = cur.x; a.origin[3 * p + 1] = cur.y; a.origin[3 * p + 2] = cur.z; }
    if (a.dir) { a.dir[3 * p + 0] = dir.x; a.dir[3 * p + 1] = dir.y; a.dir[3 * p + 2] = dir.z; }
    const V3 env = c.env_lookup();
    a.primary_color[p] = did_scatter ? make_float4(light.x, light.y, light.z, factor) : make_float4(env.x, env.y, env.z, 1.0f);
    a.info[p] = did_scatter ? 1.0f : 0.0f;
    const size_t lin = (size_t)x * H + y;
    float rec[5] = {0, 0, 0, 0, 0};
    if (did_scatter) {
        c.store_nrc_input(cur, dir, rec);
        a.infer_filter[lin / a.cfg.infer_batch_size] = 1;
    }
    float* dst = a.infer_in + 5 * lin;
#pragma unroll
    for (int k = 0; k < 5; k++) dst[k] = rec[k];
}

}  // namespace hpmdev

// ---------------------------------------------------------------------------------------------- primary rays
// Path ids are COMPACT: the pixels that reach the volume get consecutive slots (ballot + popc, one atomic per warp), so the path
// records of a frame are one dense range (the ~30 % of the pixels that are alive, not a sparse third of a frame-sized array) and the
// first delta-tracking queue is the identity.
__global__ void __launch_bounds__(128) hpm_wf_primary_kernel(const __grid_constant__ WfArgs a, uint32_t* __restrict__ q_items, uint32_t* __restrict__ q_count) {
    using namespace hpmdev;
    const uint32_t W = a.cfg.width, H = a.cfg.height;
    const uint32_t x = a.cfg.x_begin + blockIdx.x * kTileW + threadIdx.x, y = blockIdx.y * (128 / kTileW) + threadIdx.y;
    __shared__ float s_lut[256];
    Tracker c(a.sc, stage_density_lut(a.sc, s_lut));
    bool alive = false;
    V3 entry = mk(0, 0, 0), rd = mk(0, 0, 0);
    if (x < a.cfg.x_end && y < H) {
        const float u = (float)x * (1.0f / (float)W), v = (float)y * (1.0f / (float)H);
        V3 ro, exit;
        camera_ray(a.cam, u, v, &ro, &rd);
        c.init_random(u, v, a.frame_random);
        const bool sky = c.primary_ray_misses(ro, rd);
        if (!sky) c.find_entry_exit(ro, rd, &entry, &exit);
        if (!sky && !(c.sky_sdf(entry) > MAX_RAY_DISTANCE)) {
            alive = true;                  // TracePath starts at the entry point (gen_rays.comp:82-84)
        } else {
            // sky pixel: the path never starts (gen_rays.comp:78-81); nrcRayOrigin / nrcRayDir keep their cleared values
            const size_t p = (size_t)y * W + x;
            if (a.origin) { a.origin[3 * p + 0] = 0; a.origin[3 * p + 1] = 0; a.origin[3 * p + 2] = 0; }
            if (a.dir) { a.dir[3 * p + 0] = 0; a.dir[3 * p + 1] = 0; a.dir[3 * p + 2] = 0; }
            const V3 env = c.env_lookup();
            a.primary_color[p] = make_float4(env.x, env.y, env.z, 1.0f);
            a.info[p] = 0.0f;
            float* dst = a.infer_in + 5 * ((size_t)x * H + y);
#pragma unroll
            for (int k = 0; k < 5; k++) dst[k] = 0.0f;
        }
    }
    const uint32_t lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31;
    const uint32_t ballot = __ballot_sync(0xffffffffu, alive);
    if (!ballot) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(q_count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (alive) {
        const uint32_t pid = base + __popc(ballot & ((1u << lane) - 1)), n = a.st.n;
        q_items[pid] = pid;
        a.st.pixel[pid] = x | (y << 16);
        a.st.rng[pid] = c.rng;
        st3(a.st.cur, n, pid, entry); st3(a.st.dir, n, pid, rd); st3(a.st.light, n, pid, mk(0, 0, 0));
        a.st.factor[pid] = 1.0f; a.st.bounce[pid] = 0u;
    }
}

// ---------------------------------------------------------------------------------------------- paths with regeneration
constexpr uint32_t kWfChunk = 32;

struct WfRound {
    const uint32_t* in_items; const uint32_t* in_count; uint32_t* in_head;      // queue this launch drains
    uint32_t* out_items; uint32_t* out_count;                                   // queue of the spilled survivors (next launch)
    uint32_t spill_below;                                                       // 0: run every path to its end
};

__global__ void __launch_bounds__(128, 8) hpm_wf_paths_kernel(const __grid_constant__ WfArgs a, const __grid_constant__ WfRound q) {
    using namespace hpmdev;
    __shared__ float s_lut[256];
    Tracker c(a.sc, stage_density_lut(a.sc, s_lut));
    const uint32_t lane = threadIdx.x & 31, n = a.st.n;
    const uint32_t q_n = *q.in_count;
    uint32_t c_next = 0, c_end = 0;
    bool exhausted = q_n == 0;
    bool active = false, did_scatter = false;
    uint32_t pid = 0, px = 0;
    int i = 0;
    float rng = 0.0f, factor = 1.0f;
    V3 cur = mk(0, 0, 0), dir = mk(0, 0, 0), light = mk(0, 0, 0);
    uint32_t lookups = 0;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle) {
            // refill at the bounce boundary: idle lanes take the next entries of the warp's chunk in queue order
            if (c_next == c_end && !exhausted) {
                uint32_t b = 0;
                if (lane == 0) b = atomicAdd(q.in_head, kWfChunk);
                b = __shfl_sync(0xffffffffu, b, 0);
                if (b >= q_n) exhausted = true;
                else { c_next = b; c_end = min(b + kWfChunk, q_n); }
            }
            const uint32_t avail = c_end - c_next;
            if (avail) {
                const uint32_t rank = __popc(idle & ((1u << lane) - 1));
                if (!active && rank < avail) {
                    pid = q.in_items[c_next + rank];
                    px = a.st.pixel[pid];
                    rng = a.st.rng[pid];
                    cur = ld3(a.st.cur, n, pid); dir = ld3(a.st.dir, n, pid); light = ld3(a.st.light, n, pid);
                    factor = a.st.factor[pid];
                    const uint32_t bw = a.st.bounce[pid];
                    i = (int)(bw & 0xffffu); did_scatter = (bw >> 16) != 0;
                    active = true;
                }
                c_next += min(avail, (uint32_t)__popc(idle));
            } else if (exhausted) {
                if (idle == 0xffffffffu) break;
                if (32u - (uint32_t)__popc(idle) < q.spill_below) {
                    // the queue is dry and this warp is mostly empty: hand the survivors to the next launch
                    const bool keep = active;
                    const uint32_t ballot = ~idle;
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(q.out_count, __popc(ballot));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (keep) {
                        a.st.rng[pid] = rng;
                        st3(a.st.cur, n, pid, cur); st3(a.st.dir, n, pid, dir); st3(a.st.light, n, pid, light);
                        a.st.factor[pid] = factor; a.st.bounce[pid] = (uint32_t)i | (did_scatter ? 1u << 16 : 0u);
                        q.out_items[base + __popc(ballot & ((1u << lane) - 1))] = pid;
                    }
                    break;
                }
            }
        }
        bool ended = false;
        if (active) {
            // one iteration of TracePath's loop (gen_rays.comp:24-43)
            c.rng = rng; c.lookups = 0;
            bool volume_exit = false;
            cur = c.delta_track(cur, dir, &volume_exit);
            if (volume_exit) ended = true;
            else {
                did_scatter = true;
                factor *= 0.5f;
                const V3 l = c.trace_scene(cur, dir) * factor;
                light = light + l;
                dir = c.new_ray_dir(dir, true);
                if (i >= (int)a.cfg.primary_ray_length) {
                    if (c.rand_float(1.0f) >= a.cfg.primary_ray_prob || i == 128) ended = true;
                }
                i++;
            }
            rng = c.rng; lookups += c.lookups;
            if (ended) { wf_finish(a, c, px & 0xffffu, px >> 16, cur, dir, light, factor, did_scatter); active = false; }
        }
        warp_push(ended && did_scatter, (px & 0xffffu) * a.cfg.height + (px >> 16), a.active_list, a.active_count);
    }
    warp_add_u64(a.lookups, lookups);
}

}  // namespace nrchpm
