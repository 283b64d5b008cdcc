// Host side of the tracking passes + frame loop: en::NrcHpmRenderer::Render (reference src/NrcHpmRenderer.cu:299-353)
// and the command buffers it records (:1963-2087) re-expressed as one CUDA stream: clear -> gen_rays(+prep_infer_rays)
// -> prep_train_rays -> NRC inference -> NRC training -> render, with no host round trip in the compacted mode (the
// reference waits on a fence and reads the inference filter back every frame, :332-334).
#include "hpm_host.h"
#include "hpm_wavefront.cuh"
#include "nrc_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace nrchpm {

Scene::Scene(const hpm_scene_desc& d, const uint8_t* grid_host) {
    NRCHPM_REQUIRE(d.dim[0] > 0 && d.dim[1] > 0 && d.dim[2] > 0, "scene: bad grid extent");
    NRCHPM_REQUIRE(d.density_factor > 0.0f, "scene: density_factor must be > 0");
    const size_t n = (size_t)d.dim[0] * d.dim[1] * d.dim[2];
    NRCHPM_REQUIRE(n < (1ull << 32), "scene: the density grid must have fewer than 2^32 voxels");
    grid_.allocate(n);
    NRCHPM_CUDA(cudaMemcpy(grid_.ptr, grid_host, n, cudaMemcpyHostToDevice));
    dev_.grid = grid_.ptr;
    for (int i = 0; i < 3; i++) {
        dev_.dim[i] = d.dim[i]; dev_.dimf[i] = (float)d.dim[i];
        dev_.sky[i] = d.sky_size[i]; dev_.half_sky[i] = d.sky_size[i] / 2.0f; dev_.inv_sky[i] = 1.0f / d.sky_size[i];
        dev_.dl_dir[i] = d.dir_light_dir[i]; dev_.pl_pos[i] = d.point_pos[i]; dev_.pl_color[i] = d.point_color[i]; dev_.env_color[i] = d.env_color[i];
    }
    dev_.density = d.density_factor; dev_.inv_density = 1.0f / d.density_factor; dev_.g = d.g;
    dev_.dl_strength = d.dir_light_strength; dev_.pl_strength = d.point_strength; dev_.env_strength = d.env_strength;
}

Renderer::Renderer(Scene* scene, NrcCache* nrc, const hpm_render_config& cfg, cudaStream_t stream) : scene_(scene), nrc_(nrc), cfg_(cfg), stream_(stream) {
    NRCHPM_REQUIRE(scene, "renderer: null scene");
    NRCHPM_REQUIRE(cfg.width > 0 && cfg.height > 0, "renderer: bad resolution");
    NRCHPM_REQUIRE(cfg.infer_batch_size > 0, "renderer: infer_batch_size must be > 0");
    // the filter has one flag per inference batch of the cache (nrc-descriptors.glsl nrcInferFilter, NeuralRadianceCache.cu:136-144):
    // both sides take INFER_BATCH_SIZE from the same AppConfig in the reference, a mismatch would silently skip batches
    NRCHPM_REQUIRE(!nrc || cfg.compact_inference || cfg.infer_batch_size == nrc->config().infer_batch_size,
                   "renderer: infer_batch_size differs from the NeuralRadianceCache's (both come from one AppConfig)");
    if (cfg_.x_end == 0) { cfg_.x_begin = 0; cfg_.x_end = cfg_.width; }
    NRCHPM_REQUIRE(cfg_.x_begin < cfg_.x_end && cfg_.x_end <= cfg_.width, "renderer: bad column strip");
    n_pixels_ = cfg_.width * cfg_.height;
    n_train_ = cfg_.train_width * cfg_.train_height;
    NRCHPM_REQUIRE(cfg_.train_ring_size <= std::max(n_train_, 1u), "renderer: ring larger than the train pixel count (src/NrcHpmRenderer.cu:251)");
    n_filter_ = (n_pixels_ + cfg_.infer_batch_size - 1) / cfg_.infer_batch_size;
    dcfg_.width = cfg_.width; dcfg_.height = cfg_.height; dcfg_.train_width = cfg_.train_width; dcfg_.train_height = cfg_.train_height;
    dcfg_.train_x_dist = cfg_.train_x_dist; dcfg_.train_y_dist = cfg_.train_y_dist; dcfg_.train_spp = cfg_.train_spp;
    dcfg_.primary_ray_length = cfg_.primary_ray_length; dcfg_.primary_ray_prob = cfg_.primary_ray_prob;
    dcfg_.train_ring_size = cfg_.train_ring_size; dcfg_.train_ray_length = cfg_.train_ray_length; dcfg_.infer_batch_size = cfg_.infer_batch_size;
    dcfg_.x_begin = cfg_.x_begin; dcfg_.x_end = cfg_.x_end; dcfg_.train_tx0 = cfg_.train_tx0;
    blend_ = cfg_.blend != 0;

    output_.allocate(n_pixels_); primary_color_.allocate(n_pixels_); info_.allocate(n_pixels_);
    origin_.allocate((size_t)n_pixels_ * 3); dir_.allocate((size_t)n_pixels_ * 3);
    infer_in_.allocate((size_t)n_pixels_ * 5); infer_out_.allocate((size_t)n_pixels_ * 3);
    filter_.allocate(n_filter_); active_list_.allocate(n_pixels_); active_count_.allocate(1); counters_.allocate(4);
    output_.zero(); primary_color_.zero(); info_.zero(); origin_.zero(); dir_.zero(); infer_in_.zero(); infer_out_.zero();
    filter_.zero(); active_count_.zero(); counters_.zero();
    if (n_train_) {
        train_in_.allocate((size_t)n_train_ * 5); train_target_.allocate((size_t)n_train_ * 3);
        train_ray_.allocate((size_t)n_train_ * 6); train_flags_.allocate(n_train_);
        block_totals_.allocate(2 * (size_t)((n_train_ + 255) / 256));
        train_in_.zero(); train_target_.zero();
        if (cfg_.pipeline_train) {
            train_in2_.allocate((size_t)n_train_ * 5); train_target2_.allocate((size_t)n_train_ * 3);
            train_in2_.zero(); train_target2_.zero();
            // high priority: the small training grids must get SM slots as the (huge) tracking grid retires CTAs, not after it
            int prio_lo = 0, prio_hi = 0;
            NRCHPM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            NRCHPM_CUDA(cudaStreamCreateWithPriority(&train_stream_, cudaStreamNonBlocking, prio_hi));
            NRCHPM_CUDA(cudaEventCreateWithFlags(&ev_infer_done_, cudaEventDisableTiming));
            NRCHPM_CUDA(cudaEventCreateWithFlags(&ev_train_done_, cudaEventDisableTiming));
        }
        // CreateNrcTrainRingBuffer (:841-881): head = tail = 0, every ray pos (0,0,0) dir (0,0,1)
        std::vector<uint32_t> ring(2 + 6 * (size_t)n_train_, 0u);
        float* rays = reinterpret_cast<float*>(ring.data() + 2);
        for (uint32_t i = 0; i < n_train_; i++) rays[6 * (size_t)i + 5] = 1.0f;
        ring_.allocate(ring.size());
        NRCHPM_CUDA(cudaMemcpy(ring_.ptr, ring.data(), ring.size() * 4, cudaMemcpyHostToDevice));
    }
    NRCHPM_CUDA(cudaMallocHost((void**)&filter_host_, n_filter_ * sizeof(uint32_t)));
    for (auto& e : ev_) NRCHPM_CUDA(cudaEventCreate(&e));
    NRCHPM_CUDA(cudaDeviceSynchronize());
    if (nrc_) {
        NRCHPM_REQUIRE(!n_train_ || (uint64_t)nrc_->config().train_batch_size * nrc_->config().train_batch_count == n_train_,
                       "renderer: train_width*train_height must equal trainBatchCount*trainBatchSize (src/NrcHpmRenderer.cu:247-248)");
        nrc_->init(n_pixels_, infer_in_.ptr, infer_out_.ptr, train_in_.ptr, train_target_.ptr, nullptr, nullptr, stream_);
    }
}

Renderer::~Renderer() {
    if (train_stream_) { cudaStreamSynchronize(train_stream_); cudaStreamDestroy(train_stream_); cudaEventDestroy(ev_infer_done_); cudaEventDestroy(ev_train_done_); }
    if (filter_host_) cudaFreeHost(filter_host_);
    for (auto& e : ev_) if (e) cudaEventDestroy(e);
}

void Renderer::set_camera(const float m[16], const float pos[3]) {
    std::memcpy(cam_.m, m, sizeof(cam_.m));
    std::memcpy(cam_.pos, pos, sizeof(cam_.pos));
    blend_index_ = 1;                                          // SetCamera (:561-565) + image clears (:576-580)
    output_.zero(stream_); primary_color_.zero(stream_); info_.zero(stream_);
}

float Renderer::next_blend_factor() {                         // Render (:304-314)
    const float f = (float)(1.0 / (double)(float)blend_index_);
    if (blend_) blend_index_++;
    return f;
}

void Renderer::pass_gen_rays(const float fr[4]) {
    // vkCmdFillBuffer of the filter (:2002-2004); the record and info clears are folded into the kernel
    NRCHPM_CUDA(cudaMemsetAsync(filter_.ptr, 0, filter_.bytes(), stream_));
    NRCHPM_CUDA(cudaMemsetAsync(active_count_.ptr, 0, sizeof(uint32_t), stream_));
    NRCHPM_CUDA(cudaMemsetAsync(counters_.ptr, 0, sizeof(unsigned long long), stream_));
    if (use_wavefront()) { pass_gen_rays_wavefront(fr); return; }
    GenRaysArgs a{};
    a.sc = scene_->dev(); a.cam = cam_; a.cfg = dcfg_;
    a.frame_random = make_float4(fr[0], fr[1], fr[2], fr[3]);
    a.primary_color = primary_color_.ptr; a.info = info_.ptr; a.origin = origin_.ptr; a.dir = dir_.ptr;
    a.infer_in = infer_in_.ptr; a.infer_filter = filter_.ptr; a.active_list = active_list_.ptr; a.active_count = active_count_.ptr;
    a.lookups = counters_.ptr;
    const uint32_t tw = hpmdev::kTileW, th = 128 / tw;
    const dim3 block(tw, th), grid((cfg_.x_end - cfg_.x_begin + tw - 1) / tw, (cfg_.height + th - 1) / th);
    hpm_gen_rays_kernel<<<grid, block, 0, stream_>>>(a);
    check_launch("hpm_gen_rays_kernel");
}

// ---- path-regeneration form of gen_rays (hpm_wavefront.cuh): same per-pixel arithmetic and RNG streams, ended paths replaced from a queue
void Renderer::set_tracker_mode(int mode) {
    NRCHPM_REQUIRE(mode >= 0 && mode <= 2, "tracker mode: 0 automatic, 1 pixel per thread, 2 path regeneration");
    tracker_mode_ = mode;
}

bool Renderer::use_wavefront() const {
    if (cfg_.width > 65535u || cfg_.height > 65535u) return false;   // path records pack the pixel as x | y << 16
    // automatic = pixel per thread: measured faster on B200 at every configuration tried (profiles/r02_tracker_regeneration.md)
    return tracker_mode_ == 2;
}

void Renderer::pass_gen_rays_wavefront(const float fr[4]) {
    const uint32_t tw = hpmdev::kTileW, th = 128 / tw;
    const dim3 block(tw, th), grid((cfg_.x_end - cfg_.x_begin + tw - 1) / tw, (cfg_.height + th - 1) / th);
    const uint32_t n = grid.x * grid.y * 128u;                 // path slots: at most one per launched thread
    constexpr uint32_t kWords = 13;
    if (!wf_state_.ptr) {
        wf_state_.allocate((size_t)n * kWords); wf_queues_.allocate((size_t)n * 2);
        wf_counters_.allocate(2 * (kWfMaxRounds + 1));
        int dev = 0, sms = 0, occ = 0;
        NRCHPM_CUDA(cudaGetDevice(&dev));
        NRCHPM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const char* env = std::getenv("NRCHPM_WF_BLOCKS_PER_SM");
        const int want = env ? std::atoi(env) : 64;
        NRCHPM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hpm_wf_paths_kernel, 128, 0));
        wf_blocks_ = std::max(1, std::min(occ, want)) * sms;
        env = std::getenv("NRCHPM_WF_ROUNDS"); wf_rounds_ = env ? std::max(1, std::min(std::atoi(env), kWfMaxRounds)) : kWfRounds;
        env = std::getenv("NRCHPM_WF_SPILL_BELOW"); wf_spill_below_ = env ? (uint32_t)std::atoi(env) : kWfSpillBelow;
    }
    NRCHPM_CUDA(cudaMemsetAsync(wf_counters_.ptr, 0, wf_counters_.bytes(), stream_));
    WfArgs a{};
    a.sc = scene_->dev(); a.cam = cam_; a.cfg = dcfg_;
    a.frame_random = make_float4(fr[0], fr[1], fr[2], fr[3]);
    a.primary_color = primary_color_.ptr; a.info = info_.ptr; a.origin = origin_.ptr; a.dir = dir_.ptr;
    a.infer_in = infer_in_.ptr; a.infer_filter = filter_.ptr; a.active_list = active_list_.ptr; a.active_count = active_count_.ptr;
    a.lookups = counters_.ptr;
    {
        uint32_t* w = wf_state_.ptr; float* f = reinterpret_cast<float*>(w);
        size_t o = 0;
        auto take = [&](size_t words) { const size_t at = o; o += words * n; return at; };
        WfState& st = a.st;
        st.n = n;
        st.pixel = w + take(1); st.rng = f + take(1); st.cur = f + take(3); st.dir = f + take(3); st.light = f + take(3);
        st.factor = f + take(1); st.bounce = w + take(1);
    }
    uint32_t* items[2] = {wf_queues_.ptr, wf_queues_.ptr + n};
    uint32_t* ctr = wf_counters_.ptr;
    hpm_wf_primary_kernel<<<grid, block, 0, stream_>>>(a, items[0], ctr + 0);
    check_launch("hpm_wf_primary_kernel");
    // a path that can only make one bounce has nothing to regroup: one round.  Otherwise the early rounds hand their stragglers on.
    const bool short_paths = cfg_.primary_ray_length == 0 && !(cfg_.primary_ray_prob > 0.0f);
    const int rounds = short_paths ? 1 : wf_rounds_;
    for (int r = 0; r < rounds; r++) {
        WfRound q{};
        q.in_items = items[r & 1]; q.in_count = ctr + 2 * r; q.in_head = ctr + 2 * r + 1;
        q.out_items = items[(r & 1) ^ 1]; q.out_count = ctr + 2 * (r + 1);
        q.spill_below = r + 1 < rounds ? wf_spill_below_ : 0u;
        hpm_wf_paths_kernel<<<wf_blocks_, 128, 0, stream_>>>(a, q);
        check_launch("hpm_wf_paths_kernel");
    }
}

void Renderer::pass_prep_train(const float fr[4]) {
    if (!n_train_) return;
    NRCHPM_CUDA(cudaMemsetAsync(counters_.ptr + 1, 0, sizeof(unsigned long long), stream_));
    const uint32_t n_blocks = (n_train_ + 255) / 256;
    {
        TrainSelectArgs a{};
        a.cfg = dcfg_; a.info = info_.ptr; a.origin = origin_.ptr; a.dir = dir_.ptr;
        a.ring = ring_.ptr; a.train_ray = train_ray_.ptr; a.train_flags = train_flags_.ptr;
        a.block_totals = block_totals_.ptr; a.n_blocks = n_blocks;
        hpm_train_count_kernel<<<n_blocks, 256, 0, stream_>>>(a);
        check_launch("hpm_train_count_kernel");
        hpm_train_assign_kernel<<<n_blocks, 256, 0, stream_>>>(a);
        check_launch("hpm_train_assign_kernel");
    }
    {
        TrainTraceArgs a{};
        a.sc = scene_->dev(); a.cfg = dcfg_; a.frame_random = make_float4(fr[0], fr[1], fr[2], fr[3]);
        a.train_ray = train_ray_.ptr; a.train_flags = train_flags_.ptr; a.ring = ring_.ptr;
        a.block_totals = block_totals_.ptr; a.n_blocks = n_blocks;
        a.train_in = cur_train_in(); a.train_target = cur_train_target(); a.lookups = counters_.ptr + 1;
        last_train_set_ = train_set_;
        hpm_train_trace_kernel<<<(n_train_ + 127) / 128, 128, 0, stream_>>>(a);
        check_launch("hpm_train_trace_kernel");
    }
}

void Renderer::pass_composite() {
    CompositeArgs a{};
    a.cfg = dcfg_; a.primary_color = primary_color_.ptr; a.info = info_.ptr; a.infer_out = infer_out_.ptr; a.output = output_.ptr;
    a.show_nrc = cfg_.show_nrc; a.blend_factor = blend_factor_;
    const dim3 block(32, 8), grid((cfg_.x_end - cfg_.x_begin + 31) / 32, (cfg_.height + 7) / 8);
    hpm_composite_kernel<<<grid, block, 0, stream_>>>(a);
    check_launch("hpm_composite_kernel");
}

void Renderer::render(const float fr[4], bool train) {
    NRCHPM_REQUIRE(nrc_, "hpm_render needs a NeuralRadianceCache (create the renderer with one)");
    blend_factor_ = next_blend_factor();
    NRCHPM_CUDA(cudaEventRecord(ev_[0], stream_));
    pass_gen_rays(fr);
    NRCHPM_CUDA(cudaEventRecord(ev_[1], stream_));
    pass_prep_train(fr);                                       // the reference records it unconditionally (:2036-2042)
    NRCHPM_CUDA(cudaEventRecord(ev_[2], stream_));
    const bool pipelined = train_stream_ != nullptr;
    // pipelined: Train() of the previous frame has been running underneath this frame's tracking; Inference() needs its result
    if (pipelined && train_in_flight_) { NRCHPM_CUDA(cudaStreamWaitEvent(stream_, ev_train_done_, 0)); train_in_flight_ = false; }
    if (cfg_.compact_inference) {
        nrc_->inference(infer_in_.ptr, infer_out_.ptr, n_pixels_, true, active_list_.ptr, active_count_.ptr, stream_);
    } else {
        // reference behaviour: zero-filled output (:1998), host-side filter (:332-334), every record of a flagged batch
        NRCHPM_CUDA(cudaMemsetAsync(infer_out_.ptr, 0, infer_out_.bytes(), stream_));
        NRCHPM_CUDA(cudaMemcpyAsync(filter_host_, filter_.ptr, n_filter_ * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
        NRCHPM_CUDA(cudaStreamSynchronize(stream_));
        nrc_->run_inference(filter_host_);
    }
    NRCHPM_CUDA(cudaEventRecord(ev_[3], stream_));
    if (train && n_train_) {
        if (pipelined) {
            // same order of effects as the reference (Inference() of this frame, then Train(), then the next frame's Inference()):
            // the training steps start when the inference above has finished and run on their own stream, next to the tracking
            // passes of the NEXT frame (ALU-bound, no TMEM: they share the SMs with the latency-bound training kernels)
            NRCHPM_CUDA(cudaEventRecord(ev_infer_done_, stream_));
            NRCHPM_CUDA(cudaStreamWaitEvent(train_stream_, ev_infer_done_, 0));
            nrc_->run_train_on(cur_train_in(), cur_train_target(), train_stream_);
            NRCHPM_CUDA(cudaEventRecord(ev_train_done_, train_stream_));
            train_in_flight_ = true;
            train_set_ ^= 1;                                   // the next prep_train writes the other record set
        } else {
            nrc_->run_train();
        }
    }
    NRCHPM_CUDA(cudaEventRecord(ev_[4], stream_));
    pass_composite();
    NRCHPM_CUDA(cudaEventRecord(ev_[5], stream_));
    timed_ = true;
}

void Renderer::mc_render(const float fr[4], uint32_t path_length) {
    blend_factor_ = next_blend_factor();
    NRCHPM_CUDA(cudaMemsetAsync(counters_.ptr, 0, sizeof(unsigned long long), stream_));
    McArgs a{};
    a.sc = scene_->dev(); a.cam = cam_; a.cfg = dcfg_; a.frame_random = make_float4(fr[0], fr[1], fr[2], fr[3]);
    a.path_length = path_length; a.blend_factor = blend_factor_; a.output = output_.ptr; a.lookups = counters_.ptr;
    const uint32_t tw = hpmdev::kTileW, th = 128 / tw;
    const dim3 block(tw, th), grid((cfg_.x_end - cfg_.x_begin + tw - 1) / tw, (cfg_.height + th - 1) / th);
    hpm_mc_render_kernel<<<grid, block, 0, stream_>>>(a);
    check_launch("hpm_mc_render_kernel");
}

void Renderer::stage_ms(float ms[7]) {
    for (int i = 0; i < 7; i++) ms[i] = 0.0f;
    if (!timed_) return;
    NRCHPM_CUDA(cudaEventSynchronize(ev_[5]));
    ms[0] = 0.0f;                                              // clears are folded into gen_rays
    for (int i = 1; i <= 5; i++) NRCHPM_CUDA(cudaEventElapsedTime(&ms[i], ev_[i - 1], ev_[i]));
    NRCHPM_CUDA(cudaEventElapsedTime(&ms[6], ev_[0], ev_[5]));
}

void Renderer::buffer_info(int which, void** ptr, size_t* bytes) {
    void* p = nullptr; size_t b = 0;
    switch (which) {
        case HPM_BUF_OUTPUT: p = output_.ptr; b = output_.bytes(); break;
        case HPM_BUF_PRIMARY_COLOR: p = primary_color_.ptr; b = primary_color_.bytes(); break;
        case HPM_BUF_PRIMARY_INFO: p = info_.ptr; b = info_.bytes(); break;
        case HPM_BUF_NRC_ORIGIN: p = origin_.ptr; b = origin_.bytes(); break;
        case HPM_BUF_NRC_DIR: p = dir_.ptr; b = dir_.bytes(); break;
        case HPM_BUF_INFER_INPUT: p = infer_in_.ptr; b = infer_in_.bytes(); break;
        case HPM_BUF_INFER_OUTPUT: p = infer_out_.ptr; b = infer_out_.bytes(); break;
        case HPM_BUF_TRAIN_INPUT: p = last_train_in(); b = train_in_.bytes(); break;          // the set written last (not the one the next frame will write)
        case HPM_BUF_TRAIN_TARGET: p = last_train_target(); b = train_target_.bytes(); break;
        case HPM_BUF_TRAIN_RING: p = ring_.ptr; b = ring_.bytes(); break;
        case HPM_BUF_INFER_FILTER: p = filter_.ptr; b = filter_.bytes(); break;
        case HPM_BUF_COUNTERS: p = counters_.ptr; b = counters_.bytes(); break;
        default: throw Error(NRCHPM_ERR_INVALID, "unknown buffer id");
    }
    if (ptr) *ptr = p;
    if (bytes) *bytes = b;
}

void Renderer::read_buffer(int which, void* host, size_t bytes) {
    void* p; size_t b;
    buffer_info(which, &p, &b);
    NRCHPM_REQUIRE(bytes <= b, "hpm_read_buffer: size exceeds the buffer");
    NRCHPM_CUDA(cudaStreamSynchronize(stream_));
    if (train_stream_) NRCHPM_CUDA(cudaStreamSynchronize(train_stream_));     // a queued Train() still reads / the cache still writes behind the main stream
    if (which == HPM_BUF_COUNTERS) {
        // counters[2] mirrors the device-side active record count
        uint32_t cnt = 0;
        NRCHPM_CUDA(cudaMemcpy(&cnt, active_count_.ptr, 4, cudaMemcpyDeviceToHost));
        unsigned long long v = cnt;
        NRCHPM_CUDA(cudaMemcpy(counters_.ptr + 2, &v, 8, cudaMemcpyHostToDevice));
    }
    NRCHPM_CUDA(cudaMemcpy(host, p, bytes, cudaMemcpyDeviceToHost));
}

void Renderer::write_buffer(int which, const void* host, size_t bytes) {
    void* p; size_t b;
    buffer_info(which, &p, &b);
    NRCHPM_REQUIRE(bytes <= b, "hpm_write_buffer: size exceeds the buffer");
    NRCHPM_CUDA(cudaStreamSynchronize(stream_));
    NRCHPM_CUDA(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
}

}  // namespace nrchpm

// ================================================================================================ C ABI
using namespace nrchpm;
struct hpm_scene { Scene impl; hpm_scene(const hpm_scene_desc& d, const uint8_t* g) : impl(d, g) {} };
struct hpm_renderer { Renderer impl; hpm_renderer(Scene* s, NrcCache* n, const hpm_render_config& c, cudaStream_t st) : impl(s, n, c, st) {} };

extern "C" {

int hpm_scene_create(const hpm_scene_desc* desc, const uint8_t* grid_host, hpm_scene** out) {
    return guard([&] { NRCHPM_REQUIRE(desc && grid_host && out, "null argument"); *out = nullptr; *out = new hpm_scene(*desc, grid_host); });
}
int hpm_scene_destroy(hpm_scene* s) { return guard([&] { delete s; }); }
int hpm_renderer_create(hpm_scene* scene, nrc_cache* nrc, const hpm_render_config* cfg, void* stream, hpm_renderer** out) {
    return guard([&] {
        NRCHPM_REQUIRE(scene && cfg && out, "null argument");
        *out = nullptr;
        *out = new hpm_renderer(&scene->impl, nrc ? &nrc->impl : nullptr, *cfg, (cudaStream_t)stream);
    });
}
int hpm_renderer_destroy(hpm_renderer* r) { return guard([&] { delete r; }); }
int hpm_renderer_set_camera(hpm_renderer* r, const float m[16], const float pos[3]) { return guard([&] { NRCHPM_REQUIRE(r && m && pos, "null argument"); r->impl.set_camera(m, pos); }); }
int hpm_renderer_set_blend(hpm_renderer* r, int blend) { return guard([&] { NRCHPM_REQUIRE(r, "null renderer"); r->impl.set_blend(blend != 0); }); }
int hpm_render(hpm_renderer* r, const float fr[4], int train) { return guard([&] { NRCHPM_REQUIRE(r && fr, "null argument"); r->impl.render(fr, train != 0); }); }
int hpm_mc_render(hpm_renderer* r, const float fr[4], uint32_t path_length) { return guard([&] { NRCHPM_REQUIRE(r && fr, "null argument"); r->impl.mc_render(fr, path_length); }); }
int hpm_renderer_set_tracker_mode(hpm_renderer* r, int mode) { return guard([&] { NRCHPM_REQUIRE(r, "null renderer"); r->impl.set_tracker_mode(mode); }); }
int hpm_pass_gen_rays(hpm_renderer* r, const float fr[4]) { return guard([&] { NRCHPM_REQUIRE(r && fr, "null argument"); r->impl.pass_gen_rays(fr); }); }
int hpm_pass_prep_train(hpm_renderer* r, const float fr[4]) { return guard([&] { NRCHPM_REQUIRE(r && fr, "null argument"); r->impl.pass_prep_train(fr); }); }
int hpm_pass_composite(hpm_renderer* r) { return guard([&] { NRCHPM_REQUIRE(r, "null renderer"); r->impl.pass_composite(); }); }
int hpm_sync(hpm_renderer* r) { return guard([&] { NRCHPM_REQUIRE(r, "null renderer"); r->impl.sync(); }); }
int hpm_get_stage_ms(hpm_renderer* r, float ms[7]) { return guard([&] { NRCHPM_REQUIRE(r && ms, "null argument"); r->impl.stage_ms(ms); }); }
int hpm_buffer_info(hpm_renderer* r, int which, void** p, size_t* b) { return guard([&] { NRCHPM_REQUIRE(r, "null renderer"); r->impl.buffer_info(which, p, b); }); }
int hpm_read_buffer(hpm_renderer* r, int which, void* host, size_t bytes) { return guard([&] { NRCHPM_REQUIRE(r && host, "null argument"); r->impl.read_buffer(which, host, bytes); }); }
int hpm_write_buffer(hpm_renderer* r, int which, const void* host, size_t bytes) { return guard([&] { NRCHPM_REQUIRE(r && host, "null argument"); r->impl.write_buffer(which, host, bytes); }); }


// test hook: the tracker's branch-free logf against the CUDA math library on every argument the RNG can produce
int hpm_selftest_logf(uint64_t* mismatches_out) {
    return guard([&] {
        NRCHPM_REQUIRE(mismatches_out, "null argument");
        DeviceBuffer<unsigned long long> d; d.allocate(1); d.zero();
        hpm_logf_check_kernel<<<(1u << 23) / 256, 256>>>(d.ptr);
        check_launch("hpm_logf_check_kernel");
        unsigned long long h = 0;
        NRCHPM_CUDA(cudaMemcpy(&h, d.ptr, sizeof(h), cudaMemcpyDeviceToHost));
        *mismatches_out = h;
    });
}

// Reference::CompareNrc / CompareMc (src/Reference.cpp:72-171): image statistics on the device, one host read-back of 20 bytes
int hpm_compare_images(const float* d_ref_rgba, const float* d_cmp_rgba, uint32_t width, uint32_t height, hpm_compare_result* out, void* stream) {
    return guard([&] {
        NRCHPM_REQUIRE(d_ref_rgba && d_cmp_rgba && out && width && height, "null argument");
        cudaStream_t s = (cudaStream_t)stream;
        const size_t n = (size_t)width * height;
        const int blocks = (int)std::min<size_t>(kCmpBlocks, (n + kCmpThreads - 1) / kCmpThreads);
        DeviceBuffer<double> partials; partials.allocate((size_t)blocks * 5);
        DeviceBuffer<float> result; result.allocate(5);
        const float4* ref = reinterpret_cast<const float4*>(d_ref_rgba); const float4* cmp = reinterpret_cast<const float4*>(d_cmp_rgba);
        hpm_compare_pass1_kernel<<<blocks, kCmpThreads, 0, s>>>(ref, cmp, n, partials.ptr);
        check_launch("hpm_compare_pass1_kernel");
        hpm_compare_pass2_kernel<<<blocks, kCmpThreads, 0, s>>>(ref, cmp, n, partials.ptr, blocks, partials.ptr + 4 * (size_t)blocks);
        check_launch("hpm_compare_pass2_kernel");
        hpm_compare_finish_kernel<<<1, 1, 0, s>>>(partials.ptr, partials.ptr + 4 * (size_t)blocks, blocks, result.ptr);
        check_launch("hpm_compare_finish_kernel");
        float h[5];
        NRCHPM_CUDA(cudaMemcpyAsync(h, result.ptr, sizeof(h), cudaMemcpyDeviceToHost, s));
        NRCHPM_CUDA(cudaStreamSynchronize(s));
        out->mse = h[0]; out->ref_mean = h[1]; out->own_mean = h[2]; out->own_var = h[3];
        std::memcpy(&out->valid_pixel_count, &h[4], 4);
    });
}

}  // extern "C"
