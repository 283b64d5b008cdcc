// nrchpm::Scene / nrchpm::Renderer -- host objects behind the hpm_* C ABI: the CUDA-only counterpart of en::HpmScene
// (reference src/HpmScene.cpp:23-54) and en::NrcHpmRenderer (src/NrcHpmRenderer.cu:212-353, 561-642), with the
// Vulkan images / buffers of CreateNrcBuffers (:700-821), CreateNrcInferFilterBuffer (:823-839) and
// CreateNrcTrainRingBuffer (:841-881) replaced by plain device allocations.
#pragma once
#include "common.h"
#include "hpm_kernels.cuh"

namespace nrchpm {

class NrcCache;

class Scene {
public:
    Scene(const hpm_scene_desc& d, const uint8_t* grid_host);
    const SceneDev& dev() const { return dev_; }
private:
    DeviceBuffer<uint8_t> grid_;
    SceneDev dev_{};
};

class Renderer {
public:
    Renderer(Scene* scene, NrcCache* nrc, const hpm_render_config& cfg, cudaStream_t stream);
    ~Renderer();
    void set_camera(const float inv_proj_view[16], const float pos[3]);
    void set_blend(bool blend) { blend_ = blend; blend_index_ = 1; }                 // NrcHpmRenderer::SetBlend (:606-610)
    void render(const float frame_random[4], bool train);
    void mc_render(const float frame_random[4], uint32_t path_length);
    void pass_gen_rays(const float frame_random[4]);
    // 0 = automatic, 1 = one pixel per thread (hpm_gen_rays_kernel), 2 = path regeneration (hpm_wavefront.cuh); same results bit for bit
    void set_tracker_mode(int mode);
    void pass_prep_train(const float frame_random[4]);
    void pass_composite();
    void sync() { NRCHPM_CUDA(cudaStreamSynchronize(stream_)); if (train_stream_) NRCHPM_CUDA(cudaStreamSynchronize(train_stream_)); }
    void stage_ms(float ms[7]);
    void buffer_info(int which, void** ptr, size_t* bytes);
    void read_buffer(int which, void* host, size_t bytes);
    void write_buffer(int which, const void* host, size_t bytes);

private:
    float next_blend_factor();
    bool use_wavefront() const;
    void pass_gen_rays_wavefront(const float frame_random[4]);
    int tracker_mode_ = 0;
    DeviceBuffer<uint32_t> wf_state_, wf_queues_, wf_counters_;      // path records (13 words per path slot), 2 index queues, per-round counters
    int wf_blocks_ = 0, wf_rounds_ = 1;                              // persistent grid (occupancy x SM count), launches of the path kernel per frame
    uint32_t wf_spill_below_ = 0;
    Scene* scene_;
    NrcCache* nrc_;
    hpm_render_config cfg_;
    RenderCfgDev dcfg_{};
    CameraDev cam_{};
    cudaStream_t stream_;
    bool blend_ = false;
    uint32_t blend_index_ = 1;
    float blend_factor_ = 1.0f;
    uint32_t n_pixels_ = 0, n_train_ = 0, n_filter_ = 0;
    DeviceBuffer<float4> output_, primary_color_;
    DeviceBuffer<float> info_, origin_, dir_, infer_in_, infer_out_, train_in_, train_target_, train_ray_;
    // pipelined training (cfg.pipeline_train): Train() of frame N runs on its own stream while frame N+1 is tracked; the train
    // records are double-buffered so that prep_train(N+1) does not overwrite what Train(N) still reads
    DeviceBuffer<float> train_in2_, train_target2_;
    cudaStream_t train_stream_ = nullptr;
    cudaEvent_t ev_infer_done_ = nullptr, ev_train_done_ = nullptr;
    bool train_in_flight_ = false;
    int train_set_ = 0;              // record set the NEXT prep_train pass writes (pipelined training double-buffers the records)
    int last_train_set_ = 0;         // record set the LAST prep_train pass wrote: what hpm_buffer_info / hpm_read_buffer expose
    float* cur_train_in() { return train_set_ ? train_in2_.ptr : train_in_.ptr; }
    float* cur_train_target() { return train_set_ ? train_target2_.ptr : train_target_.ptr; }
    float* last_train_in() { return last_train_set_ ? train_in2_.ptr : train_in_.ptr; }
    float* last_train_target() { return last_train_set_ ? train_target2_.ptr : train_target_.ptr; }
    DeviceBuffer<uint32_t> ring_, filter_, active_list_, active_count_, train_flags_, block_totals_;
    DeviceBuffer<unsigned long long> counters_;
    uint32_t* filter_host_ = nullptr;     // pinned
    cudaEvent_t ev_[7]{};
    bool timed_ = false;
};

}  // namespace nrchpm
