// nrchpm::NrcCache -- host object behind the nrc_* C ABI; the B200-native counterpart of en::NeuralRadianceCache
// (reference include/engine/graphics/NeuralRadianceCache.hpp:10-64) fused with the tcnn::Trainer state it owns
// (tiny-cuda-nn/include/tiny-cuda-nn/trainer.h:322-336: fp32 master, fp16 working weights, fp16 gradients;
// optimizers/adam.h, optimizers/ema.h state).
#pragma once
#include <cuda_fp16.h>
#include <string>
#include <utility>
#include <vector>
#include "common.h"
#include "nrc_types.h"

namespace nrchpm {

struct NrcConfig {
    int pos_enc = 0, dir_enc = 0;                 // src/AppConfig.cpp:16-73 presets
    int n_neurons = 64, n_hidden_layers = 5;
    bool oneblob_soa_bug = true;
    int n_levels = 16, n_features = 2, log2_hashmap_size = 19, base_resolution = 16;
    float per_level_scale = 2.0f;
    int n_freq_pos = 12, n_freq_dir = 4, n_bins = 4;
    float learning_rate = 1e-3f, ema_decay = 0.99f, beta1 = 0.9f, beta2 = 0.999f, epsilon = 1e-8f, l2_reg = 1e-8f;   // adam.h:313-317
    float loss_scale = 128.0f;                    // common.h:232 (fp16 network precision)
    uint32_t infer_batch_size = 1u << 21, train_batch_size = 1u << 14, train_batch_count = 4;
    static NrcConfig from_json(const std::string& text);
};

constexpr uint32_t kMaxDwChunks = 256;      // weight-gradient partials: 32 batch chunks up to 1024 tiles (the reference's 2^14 batches), up to 256 beyond

struct OptArgs;          // nrc_kernels.cuh

class NrcCache {
public:
    NrcCache(const NrcConfig& cfg, uint64_t seed);
    ~NrcCache();
    const NrcConfig& config() const { return cfg_; }
    const EncParams& enc() const { return enc_; }
    uint64_t n_params() const { return n_params_; }
    uint64_t n_mlp() const { return n_mlp_; }
    bool initialised() const { return initialised_; }
    size_t infer_batch_count() const { return infer_batches_.size(); }

    void init(uint32_t infer_count, float* in, float* out, float* train_in, float* train_target, cudaExternalSemaphore_t start,
              cudaExternalSemaphore_t finished, cudaStream_t stream);
    void infer_and_train(const uint32_t* filter_host, bool train);
    void run_inference(const uint32_t* filter_host);
    void run_train();
    void wait_start();
    void signal_finished();
    float loss();

    void encode(const float* d_in, uint32_t n, bool use_ema, void* d_out_half, cudaStream_t s);
    void inference(const float* d_in, float* d_out, uint32_t n, bool use_ema, const uint32_t* d_indices, const uint32_t* d_count, cudaStream_t s);
    void inference_set(int param_set, const float* d_in, float* d_out, uint32_t n, const uint32_t* d_indices, const uint32_t* d_count, cudaStream_t s);
    void snapshot_params(bool use_ema, cudaStream_t s);
    void set_inference_cta_limit(uint32_t max_ctas) { infer_max_ctas_ = max_ctas; }
    void training_step(const float* d_in, const float* d_target, uint32_t B, bool run_optimizer, cudaStream_t s);
    void optimizer_step(cudaStream_t s);
    void inference_host(const float* h_in, float* h_out, uint32_t n, bool use_ema);
    void training_step_host(const float* h_in, const float* h_tgt, uint32_t B, float* loss_out);
    void infer_and_train_host(const float* h_in, float* h_out, uint32_t n, const float* h_tin, const float* h_tgt, uint32_t B, uint32_t n_batches,
                              bool use_ema, float* loss_out);
    void run_train_on(const float* d_in, const float* d_target, cudaStream_t s);

    void get_params(int which, float* out);
    void set_params_fp32(const float* host_master);
    void set_ema(const float* host_ema);
    void gradient_buffers(float** mlp, void** enc);
    // data-parallel exchange over peer memory (cudaIpc): export this rank's handles, import every rank's, exchange one step
    static constexpr size_t kPeerHandleBytes = 3 * 64;
    void peer_export(uint8_t* out);
    void peer_setup(int rank, int world, const uint8_t* handles);
    void peer_exchange(cudaStream_t s);
    void peer_publish(cudaStream_t s);
    void peer_slice(uint64_t& begin, uint64_t& end) const;
    void fill_peer_args(struct PeerArgs& a);
    void last_step_tensor(int which, float* host_out);
    uint32_t train_profile(long long* host_out, uint32_t max_ctas);
    uint32_t read_timeline(unsigned long long* host_out, uint32_t max_slots);
    void keep_dx(bool k) { keep_dx_ = k; }
    cudaStream_t stream() const { return stream_; }

private:
    void derive();
    void init_params(uint64_t seed);
    void setup_kernels();
    void scatter_grid_field(int field, const float* d_src);
    void ensure_train_scratch(uint32_t B);
    bool fused_training_fits() const;
    bool wide() const { return cfg_.n_neurons == 128; }      // 128-neuron network: nrc_wide_kernels.cuh
    int wide_wgs_ = 1, wide_ws_slots_ = 0;
    void training_step_three_kernels(const float* d_in, const float* d_target, uint32_t B, cudaStream_t s);
    void launch_shape(uint32_t tiles, uint32_t& grid, uint32_t& threads) const;
    void ensure_pipeline(uint32_t n_chunks);
    void trace_mark(const char* name, cudaStream_t s);
    void launch_ema(const OptArgs& a);
    void queue_inference_copies_in(const float* h_in, const std::vector<uint32_t>& chunk_offsets);
    void queue_inference_pipeline(const __half* params, const float* h_in, float* h_out, const std::vector<uint32_t>& chunk_offsets, bool copies_queued);
    void inference_with(const __half* params, const float* d_in, float* d_out, uint32_t n, const uint32_t* d_indices, const uint32_t* d_count, cudaStream_t s, uint32_t max_ctas);
    void infer_and_train_overlapped();

    NrcConfig cfg_;
    EncParams enc_;
    size_t n_mlp_ = 0, n_grid_ = 0, n_params_ = 0;
    int sm_count_ = 148;
    size_t infer_smem_level_bytes() const;
    int infer_groups_ = 0, train_groups_ = 1;      // 256-thread groups per CTA of the two-threads-per-record kernels (0: one-thread kernels)
    DeviceBuffer<float> master_, m1_, m2_, loss_dev_, loss_partials_, dw_partials_, mlp_grad_f32_;
    // fp16 working weights and fp16 gradients live in ONE allocation [grad16 | w16]: the two vectors every training kernel gathers from /
    // scatters into.  Together 57 MB for the default preset -- they are pinned in the set-aside (persisting) part of the 126 MB L2 with an
    // access-policy window on every training launch, so the 224 MB of optimizer state that stream through L2 every step cannot evict them.
    struct HalfView { __half* ptr = nullptr; size_t count = 0; size_t bytes() const { return count * sizeof(__half); } };
    DeviceBuffer<__half> hot_;
    HalfView w16_, grad16_;
    size_t l2_window_bytes_ = 0; float l2_hit_ratio_ = 1.0f;
    template <class K, class A> void launch_hot(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const A& args, bool pdl = false);
    DeviceBuffer<__half> ema16_, x16_, acts_, dacts_, out16_, dout16_, dx16_;
    DeviceBuffer<uint32_t> steps_;
    DeviceBuffer<GridAdamState> grid_state_;       // Adam state of the encoding parameters, one 32-byte record per entry
    DeviceBuffer<float> host_in_, host_out_, host_tin_, host_tgt_;
    uint32_t scratch_batch_ = 0, last_batch_ = 0, dw_chunks_ = 0, current_step_ = 0;
    const float* dw_source_ = nullptr;
    bool grid_grad_dirty_ = false, grads_pending_ = false, keep_dx_ = true, loss_valid_ = true, initialised_ = false;
    float loss_host_ = 0;
    cudaStream_t train_stream_last_ = nullptr;       // stream of the last training step: loss(), gradient_buffers() order themselves behind it
    bool train_fused_ = true; int train_tpr_ = 2;
    int infer_ws_ = 1;                               // warp-specialised inference kernel (0: tile-per-warpgroup kernel)
    DeviceBuffer<unsigned int> train_done_;
    DeviceBuffer<long long> train_prof_;
    DeviceBuffer<unsigned long long> timeline_;
    static constexpr uint32_t kTimelineSlots = 256;
    uint32_t tl_seq_ = 0;
    unsigned long long* timeline_slot();
    void reset_timeline();
    float* loss_pinned_ = nullptr;
    cudaStream_t ema_stream_ = nullptr;              // dense EMA pass of the encoding weights, underneath the next training step
    cudaEvent_t adam_done_ = nullptr, ema_done_ = nullptr;
    bool ema_in_flight_ = false;
    bool pdl_enabled_ = true;                        // programmatic dependent launch of the training chain (NRCHPM_PDL=0 disables)
    void wait_ema(cudaStream_t s);
    // en::NeuralRadianceCache::Init state
    uint32_t infer_count_ = 0;
    float *infer_in_ = nullptr, *infer_out_ = nullptr, *train_in_ = nullptr, *train_target_ = nullptr;
    cudaExternalSemaphore_t start_sem_ = nullptr, finished_sem_ = nullptr;
    cudaStream_t stream_ = nullptr;
    std::vector<std::pair<uint32_t, uint32_t>> infer_batches_;   // (first record, count)
    cudaStream_t copy_in_stream_ = nullptr, copy_out_stream_ = nullptr, compute_stream_ = nullptr;   // host-buffer pipeline
    cudaStream_t train_stream_ = nullptr;            // InferAndTrain on host buffers: training overlaps the inference pipeline
    DeviceBuffer<__half> infer_snapshot_;            // ... which then reads a snapshot of the pre-training parameters
    bool snapshot_valid_ = false;
    // data-parallel replicas: InferAndTrain with the frame's inference underneath the gradient exchanges (infer_and_train_overlapped)
    static constexpr uint32_t kMaxOverlapBatches = 16;
    cudaStream_t ov_inf_stream_ = nullptr, ov_tr_stream_ = nullptr;
    cudaEvent_t ov_ev_[5 + 2 * kMaxOverlapBatches] = {};
    cudaEvent_t ema_gate_ = nullptr;
    bool overlap_schedule_ = true, peer_one_cta_per_sm_ = false; double overlap_head_ = 0.0; uint32_t peer_ctas_ = 0, overlap_infer_sms_ = 112;
    int peer_rank_ = -1, peer_world_ = 0;
    uint32_t peer_token_ = 0;
    bool peer_fused_ = true;                         // reduce-scatter + Adam + weight all-gather as one kernel (nrc_peer_adam_kernel)
    bool peer_sharded_step_ = false;                 // peer_exchange ran: the pending optimizer step covers this rank's slice and ends with the weight all-gather
    float grad_scale_ = 1.0f;                        // ranks whose gradients were summed into the buffers (the optimizer divides)
    void* peer_grad_[8] = {}; void* peer_mlp_[8] = {}; void* peer_flags_[8] = {};
    DeviceBuffer<float> mlp_sum_;
    DeviceBuffer<uint32_t> peer_flag_words_;
    DeviceBuffer<unsigned int> peer_done_;
    uint32_t infer_max_ctas_ = 0;                    // 0: the full persistent grid (2 CTAs per SM)
    std::vector<cudaEvent_t> pipe_events_;
    bool trace_on_ = false;                              // NRCHPM_E2E_TRACE: timing events of one nrc_infer_and_train_host call
    std::vector<std::pair<std::string, cudaEvent_t>> trace_;
};

}  // namespace nrchpm

// the opaque handle of the C ABI
struct nrc_cache {
    nrchpm::NrcCache impl;
    nrc_cache(const nrchpm::NrcConfig& c, uint64_t seed) : impl(c, seed) {}
};
