/* nrc_hpm_b200.h -- C ABI of the B200-native NRC-HPM hot path.
 *
 * Drop-in boundary for the two reference classes that own the per-frame hot path
 *   en::NeuralRadianceCache   (reference include/engine/graphics/NeuralRadianceCache.hpp:10-64,
 *                              src/NeuralRadianceCache.cu:11-178)
 *   en::NrcHpmRenderer        (reference include/engine/graphics/renderer/NrcHpmRenderer.hpp:13-41,
 *                              src/NrcHpmRenderer.cu:212-353) and the five GLSL passes it records
 *                              (data/shader/nrc/{clear,gen_rays,prep_infer_rays,prep_train_rays,render}.comp)
 * plus the McHpmRenderer path tracer that shares the tracking code (data/shader/mc/render.comp:7-84).
 *
 * Plain pointers and sizes only; no C++ / torch / Vulkan types.  CUDA handles cross the ABI as void*:
 *   stream      = cudaStream_t            (NULL = legacy default stream, as the reference uses)
 *   semaphore   = cudaExternalSemaphore_t (NULL = no wait / no signal)
 * Every function returns NRCHPM_OK or an error code; nrchpm_last_error() holds the message of the
 * last failure on the calling thread (the reference throws std::runtime_error instead, src/Log.cpp:16-21).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with NRCHPM_ERR_CUDA.
 */
#ifndef NRC_HPM_B200_H
#define NRC_HPM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRCHPM_OK 0
#define NRCHPM_ERR_INVALID 1      /* bad argument / bad config                      */
#define NRCHPM_ERR_CUDA 2         /* CUDA runtime error (message in last_error)      */
#define NRCHPM_ERR_UNSUPPORTED 3  /* valid for the reference, not implemented here   */

const char* nrchpm_last_error(void);
int nrchpm_version(void);
/* number of kernels launched by this library on the calling process so far (bench.py "gpu_launches") */
uint64_t nrchpm_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * NeuralRadianceCache
 * ---------------------------------------------------------------------------------------------- */
typedef struct nrc_cache nrc_cache;

/* en::NeuralRadianceCache::NeuralRadianceCache(const AppConfig&) (src/NeuralRadianceCache.cu:11-40).
 * config_json is the tiny-cuda-nn model JSON the reference builds there: {"loss":{"otype":"RelativeL2Luminance"},
 * "optimizer":{"otype":"EMA","decay":d,"nested":{"otype":"Adam","learning_rate":lr}},
 * "encoding":{"otype":"Composite","nested":[pos,dir]}, "network":{"otype":"FullyFusedMLP","activation":"ReLU",
 * "output_activation":"None","n_neurons":W,"n_hidden_layers":H}} -- W = 64 or 128 (AppConfig nnWidth, src/AppConfig.cpp:169;
 * the 128-neuron network takes at most 6 hidden layers: its weights stay resident in shared memory) -- plus the three AppConfig batch fields as
 * optional top-level keys "infer_batch_size", "train_batch_size", "train_batch_count" (defaults 2^21, 2^14, 4)
 * and "compat":{"oneblob_soa_bug":true|false} (default true = bug-for-bug tcnn behaviour).
 * seed: tcnn Trainer seed (1337 in the reference, trainer.h:50-57). */
int nrc_create(const char* config_json, uint64_t seed, nrc_cache** out);
/* en::NeuralRadianceCache::Destroy (src/NeuralRadianceCache.cu:105-112) */
int nrc_destroy(nrc_cache* c);

/* en::NeuralRadianceCache::Init (src/NeuralRadianceCache.cu:42-95).  All four buffers are DEVICE pointers owned
 * by the caller: infer_in float[infer_count][5], infer_out float[infer_count][3], train_in
 * float[batch_count*batch_size][5], train_target float[...][3].  infer_count must be a multiple of 16. */
int nrc_init(nrc_cache* c, uint32_t infer_count, float* d_infer_in, float* d_infer_out, float* d_train_in,
             float* d_train_target, void* start_semaphore, void* finished_semaphore, void* stream);
/* Vulkan <-> CUDA interop of the reference, Linux flavour.  The reference exports its four NRC record buffers with
 * VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT and imports + maps them in NrcHpmRenderer::CreateNrcBuffers (src/NrcHpmRenderer.cu:700-821);
 * its two semaphores are exported with vkGetSemaphoreFdKHR and imported in CreateSyncObjects (:644-690).  These helpers are that import
 * step: the returned device pointer / semaphore go straight into nrc_init.  The fd is owned by CUDA after a successful import (do not
 * close it).  UNTESTED against a Vulkan device: this image has none; the error paths are under test. */
typedef struct nrchpm_external_buffer nrchpm_external_buffer;
int nrchpm_import_external_buffer(int opaque_fd, size_t bytes, nrchpm_external_buffer** out, void** d_ptr_out);
int nrchpm_release_external_buffer(nrchpm_external_buffer* b);              /* NrcHpmRenderer::Destroy (:355-...) cudaDestroyExternalMemory */
int nrchpm_import_external_semaphore(int opaque_fd, void** cuda_external_semaphore_out);
int nrchpm_release_external_semaphore(void* cuda_external_semaphore);

/* en::NeuralRadianceCache::InferAndTrain / Inference / Train (src/NeuralRadianceCache.cu:97-156).
 * infer_filter is a HOST pointer, one uint32 per inference batch (NULL = run every batch). */
int nrc_infer_and_train(nrc_cache* c, const uint32_t* infer_filter_host, int train);
int nrc_inference(nrc_cache* c, const uint32_t* infer_filter_host);
int nrc_train(nrc_cache* c);
/* en::NeuralRadianceCache::GetLoss (NeuralRadianceCache.hpp:38): loss of the last training batch.  The reference
 * blocks inside Train(); here the value is fetched (stream sync) only when asked for. */
int nrc_get_loss(nrc_cache* c, float* loss);
/* GetInferBatchCount / GetTrainBatchCount / GetInferBatchSize / GetTrainBatchSize (NeuralRadianceCache.hpp:40-46) */
size_t nrc_get_infer_batch_count(const nrc_cache* c);
size_t nrc_get_train_batch_count(const nrc_cache* c);
uint32_t nrc_get_infer_batch_size(const nrc_cache* c);
uint32_t nrc_get_train_batch_size(const nrc_cache* c);

/* ---- tcnn-level surface (tcnn::cpp::Module, tiny-cuda-nn/include/tiny-cuda-nn/cpp_api.h:88-117) ---- */
uint64_t nrc_n_params(const nrc_cache* c);        /* Module::n_params            */
uint64_t nrc_n_mlp_params(const nrc_cache* c);    /* network part of the [network | encoding] parameter vector */
uint32_t nrc_input_width(const nrc_cache* c);     /* padded encoding width (48 for the default preset) */
/* which: 0 fp32 master, 1 fp16 working, 2 fp16 EMA (inference weights), 3 fp16 gradient of the last step,
 *        4 Adam first moment, 5 Adam second moment, 6 per-parameter step count.  host_out: float[n_params]. */
int nrc_get_params(nrc_cache* c, int which, float* host_out);
int nrc_set_params_fp32(nrc_cache* c, const float* host_master);   /* Trainer::set_params_full_precision */
int nrc_set_ema(nrc_cache* c, const float* host_ema);
/* Device pointers for data-parallel training: the fp32 MLP gradient accumulator (n_mlp floats) and the fp16
 * encoding gradient (n_params - n_mlp halfs), valid between nrc_training_step(run_optimizer=0) and
 * nrc_optimizer_step. */
int nrc_gradient_buffers(nrc_cache* c, float** d_mlp_grad_f32, void** d_enc_grad_f16);

/* Data-parallel training without a library collective: the gradients of nrc_training_step(run_optimizer=0) are summed
 * across the ranks of ONE node by a kernel of this library that loads and stores peer HBM over NVLink (buffers shared with
 * cudaIpc).  nrc_peer_export fills NRC_PEER_HANDLE_BYTES bytes for this rank; the caller gathers them from every rank (any
 * transport; rank-major) and passes all of them to nrc_peer_setup; nrc_peer_exchange then replaces the all-reduce of
 * nrc_gradient_buffers: afterwards nrc_optimizer_step applies the MEAN gradient.  Every rank must call it once per step.
 * After nrc_peer_setup, nrc_train / nrc_infer_and_train / hpm_render(train=1) exchange after every training step by themselves
 * (all ranks must then render in lockstep). */
#define NRC_PEER_HANDLE_BYTES 192
int nrc_peer_export(nrc_cache* c, uint8_t* handles_out);
int nrc_peer_setup(nrc_cache* c, int rank, int world, const uint8_t* all_handles);
int nrc_peer_exchange(nrc_cache* c, void* stream);

/* Encoding only: d_in float[n][5] -> d_out __half[n][input_width] (row per record). */
int nrc_encode_batch(nrc_cache* c, const float* d_in, uint32_t n, int use_ema, void* d_out_half, void* stream);
/* Network::inference on n records (any n > 0; tcnn requires a multiple of 256, object.h:150). use_ema=1 is what the
 * reference's Inference() does. */
int nrc_inference_batch(nrc_cache* c, const float* d_in, float* d_out, uint32_t n, int use_ema, void* stream);
/* use_ema = 2 in nrc_inference_batch / nrc_inference_indexed evaluates the parameters captured by the last
 * nrc_snapshot_params (a device-to-device copy of the EMA or the working weights, stream-ordered on `stream`).  A caller that
 * overlaps Inference() with Train() -- e.g. to hide the gradient all-reduce of a multi-GPU training step -- snapshots first,
 * so that the cache is still evaluated with the parameters of the previous frame (src/NeuralRadianceCache.cu:97-156 order). */
int nrc_snapshot_params(nrc_cache* c, int use_ema, void* stream);
/* Caps the persistent grid of the following inference launches (0 = no cap: two CTAs per SM).  With one CTA per SM a
 * co-running kernel -- the NCCL all-reduce of a data-parallel training step -- finds registers and shared memory on every SM. */
int nrc_set_inference_cta_limit(nrc_cache* c, uint32_t max_ctas);
/* Compacted inference: only records d_indices[0 .. *d_count) are evaluated (read from d_in[idx], written to
 * d_out[idx]); d_count is a DEVICE counter so no host sync is needed.  max_n bounds the launch. */
int nrc_inference_indexed(nrc_cache* c, const float* d_in, float* d_out, const uint32_t* d_indices,
                          const uint32_t* d_count, uint32_t max_n, int use_ema, void* stream);
/* Trainer::training_step on one batch (trainer.h:163-190): forward, RelativeL2Luminance loss, backward, and -- when
 * run_optimizer != 0 -- Adam + EMA.  batch must be a multiple of 128. */
int nrc_training_step(nrc_cache* c, const float* d_in, const float* d_target, uint32_t batch, int run_optimizer,
                      void* stream);
int nrc_optimizer_step(nrc_cache* c, void* stream);
/* test hooks: tensors of the last training step.  which: 0 padded output __half[B][16], 1 dL/doutput __half[B][16],
 * 2 dL/dinput __half[B][input_width] (only when the position encoding has parameters). host_out receives floats. */
int nrc_last_step_tensor(nrc_cache* c, int which, float* host_out);
/* development aid: clock64 stamps of the last fused training launch, long long[ctas][16] (phase boundaries of thread 0 of every
 * CTA); only recorded when the process runs with NRCHPM_TRAIN_PROF=1.  Returns the number of CTAs written (0: not recording). */
uint32_t nrc_debug_train_profile(nrc_cache* c, long long* host_out, uint32_t max_ctas);
/* development aid (NRCHPM_TRAIN_PROF=1): device-side timeline of the training kernels launched since the last call, in launch order
 * (fused step, optimizer, EMA pass, ...): host_out[2k] = earliest CTA start, host_out[2k+1] = latest CTA end, %globaltimer ns. */
uint32_t nrc_debug_timeline(nrc_cache* c, unsigned long long* host_out, uint32_t max_slots);

/* Host-buffer entry points (pageable or pinned host memory; H2D and D2H copies happen inside the call). */
int nrc_inference_host(nrc_cache* c, const float* h_in, float* h_out, uint32_t n, int use_ema);
int nrc_training_step_host(nrc_cache* c, const float* h_in, const float* h_target, uint32_t batch, float* loss_out);
/* en::NeuralRadianceCache::InferAndTrain (src/NeuralRadianceCache.cu:97-156) on HOST buffers, one call per frame: inference
 * of h_infer_in float[n_infer][5] -> h_infer_out float[n_infer][3] with the weights as they are on entry (EMA when use_ema),
 * then n_batches training steps of `batch` records each from h_train_in float[n_batches*batch][5] / h_train_target
 * float[...][3]; *loss_out = loss of the last batch (GetLoss).  H2D copies, kernels and D2H copies are pipelined over three
 * streams and the host waits once at the end.  n_infer == 0 or n_batches == 0 skip the respective half. */
int nrc_infer_and_train_host(nrc_cache* c, const float* h_infer_in, float* h_infer_out, uint32_t n_infer, const float* h_train_in,
                             const float* h_train_target, uint32_t batch, uint32_t n_batches, int use_ema, float* loss_out);

/* ------------------------------------------------------------------------------------------------
 * Scene + NrcHpmRenderer / McHpmRenderer
 * ---------------------------------------------------------------------------------------------- */
typedef struct hpm_scene hpm_scene;
typedef struct hpm_renderer hpm_renderer;

/* HpmScene + VolumeData + lights (src/HpmScene.cpp:23-54, src/AppConfig.cpp:93-150) and the specialization constants
 * that describe them (src/NrcHpmRenderer.cu:908-936). */
typedef struct hpm_scene_desc {
    int32_t dim[3];            /* density grid extent W,H,D; linear index i + W*j + W*H*k (Texture3D.cpp:107) */
    float sky_size[3];         /* normalize(extent) * 107.5 (NrcHpmRenderer.cu:910-912) */
    float density_factor;      /* VOLUME_DENSITY_FACTOR */
    float g;                   /* VOLUME_G (HpmScene.cpp:45) */
    float dir_light_dir[3];    /* DirLight::VecFromAngles (DirLight.cpp:5-14) */
    float dir_light_strength;
    float point_pos[3];
    float point_strength;
    float point_color[3];
    float env_strength;        /* HDR_ENV_MAP_STRENGTH */
    float env_color[3];        /* constant-colour environment map (1x1 texture) */
} hpm_scene_desc;

/* grid_host: dense 8-bit density (Texture3D::FromVDB semantics, one byte per voxel), copied to the device. */
int hpm_scene_create(const hpm_scene_desc* desc, const uint8_t* grid_host, hpm_scene** out);
int hpm_scene_destroy(hpm_scene* s);

/* Constructor arguments of NrcHpmRenderer + the AppConfig fields its specialization constants use
 * (src/NrcHpmRenderer.cu:212-297, 908-1061; data/shader/include/nrc-constants.glsl:1-22). */
typedef struct hpm_render_config {
    uint32_t width, height;
    uint32_t train_width, train_height, train_x_dist, train_y_dist;   /* CalcTrainSubset (:612-642) */
    uint32_t train_spp;
    uint32_t primary_ray_length;
    float primary_ray_prob;
    uint32_t train_ring_size;
    uint32_t train_ray_length;
    uint32_t infer_batch_size;
    uint32_t blend;            /* progressive blending (ctor arg `blend`) */
    uint32_t show_nrc;         /* uniform showNrc (NrcHpmRenderer.hpp:70-75) */
    uint32_t compact_inference;/* 1: evaluate the cache only at scattered pixels (warp-compacted index list);
                                  0: reference behaviour, every record of every batch whose filter flag is set */
    /* screen partition for multi-GPU runs: this renderer owns pixel columns [x_begin, x_end) */
    uint32_t x_begin, x_end;
    /* ... and the train-pixel lattice columns [train_tx0, train_tx0 + train_width) of the frame's lattice (train pixel (tx,ty)
     * sits at render pixel ((train_tx0+tx)*train_x_dist, ty*train_y_dist) and seeds its RNG with the frame-wide lattice column) */
    uint32_t train_tx0;
    /* 1: Train() of frame N runs on its own stream underneath the tracking passes of frame N+1 (double-buffered train
     * records); the order of effects stays the reference's: Inference(N), Train(N), Inference(N+1).  hpm_render returns with
     * the training in flight; hpm_sync / nrc_get_loss wait for it.  HPM_BUF_TRAIN_INPUT / _TARGET then name the record set the
     * NEXT frame will write.  0: everything on one stream, in the reference's order. */
    uint32_t pipeline_train;
} hpm_render_config;

/* nrc may be NULL (pass-level use / McHpmRenderer only).  stream: cudaStream_t or NULL. */
int hpm_renderer_create(hpm_scene* scene, nrc_cache* nrc, const hpm_render_config* cfg, void* stream, hpm_renderer** out);
int hpm_renderer_destroy(hpm_renderer* r);
/* NrcHpmRenderer::SetCamera (:561-…): invProjView as glm stores it (column-major float[16]) and the camera position
 * (Camera::UpdateUniformBuffer, src/Camera.cpp:164-174). */
int hpm_renderer_set_camera(hpm_renderer* r, const float inv_proj_view[16], const float cam_pos[3]);
int hpm_renderer_set_blend(hpm_renderer* r, int blend);       /* NrcHpmRenderer::SetBlend */
/* NrcHpmRenderer::Render(queue, train) (:299-353): clear, gen_rays, prep_infer_rays, prep_train_rays, NRC
 * InferAndTrain, render.  frame_random is the vec4 the reference draws with glm::linearRand (:305-308). */
int hpm_render(hpm_renderer* r, const float frame_random[4], int train);
/* McHpmRenderer::Render (src/McHpmRenderer.cpp:121-151, data/shader/mc/render.comp:7-84) */
int hpm_mc_render(hpm_renderer* r, const float frame_random[4], uint32_t path_length);
/* Schedule of the gen_rays pass: 0 automatic (default), 1 one pixel per thread for the whole path (the shader's own shape,
 * data/shader/nrc/gen_rays.comp:53-101), 2 path regeneration: primary rays in one coherent launch, then persistent warps in which a
 * lane whose path has ended takes the next path from a queue (TracePath, gen_rays.comp:7-51, one bounce per iteration).  Every output
 * is bit-identical between the modes except the ORDER of the compacted record list.  Automatic currently means 1: measured faster
 * on B200 (profiles/r02_tracker_regeneration.md). */
int hpm_renderer_set_tracker_mode(hpm_renderer* r, int mode);
/* individual passes, for parity tests and profiling */
int hpm_pass_gen_rays(hpm_renderer* r, const float frame_random[4]);      /* clear + gen_rays + prep_infer_rays */
int hpm_pass_prep_train(hpm_renderer* r, const float frame_random[4]);    /* clear.comp + prep_train_rays      */
int hpm_pass_composite(hpm_renderer* r);                                   /* render.comp                        */
int hpm_sync(hpm_renderer* r);
/* NrcHpmRenderer::GetFrameTimeMS / EvaluateTimestampQueries (:495-559): ms of the last frame per stage:
 * 0 clear, 1 gen_rays(+prep_infer), 2 prep_train, 3 nrc inference, 4 nrc training, 5 render, 6 total. */
int hpm_get_stage_ms(hpm_renderer* r, float ms[7]);

enum {
    HPM_BUF_OUTPUT = 0,        /* float[W*H][4], pixel (x,y) at y*W+x (outputImage)                     */
    HPM_BUF_PRIMARY_COLOR = 1, /* float[W*H][4] rgb + throughput (primaryRayColorImage)                 */
    HPM_BUF_PRIMARY_INFO = 2,  /* float[W*H] didScatter 0/1 (primaryRayInfoImage.x)                     */
    HPM_BUF_NRC_ORIGIN = 3,    /* float[W*H][3] (nrcRayOriginImage)                                     */
    HPM_BUF_NRC_DIR = 4,       /* float[W*H][3] (nrcRayDirImage)                                        */
    HPM_BUF_INFER_INPUT = 5,   /* float[W*H][5] at x*H+y (nrcInferInput)                                */
    HPM_BUF_INFER_OUTPUT = 6,  /* float[W*H][3] at x*H+y (nrcInferOutput)                               */
    HPM_BUF_TRAIN_INPUT = 7,   /* float[T][5] at y*trainW+x                                             */
    HPM_BUF_TRAIN_TARGET = 8,  /* float[T][3]                                                           */
    HPM_BUF_TRAIN_RING = 9,    /* uint32 head, tail, then float[ring][6] (nrcTrainRing)                 */
    HPM_BUF_INFER_FILTER = 10, /* uint32 per inference batch                                            */
    HPM_BUF_COUNTERS = 11      /* uint64[4]: density lookups gen_rays, lookups prep_train, active records, reserved */
};
/* size in bytes of a buffer / device pointer / copy to or from host (stream-synchronous) */
int hpm_buffer_info(hpm_renderer* r, int which, void** d_ptr, size_t* bytes);
int hpm_read_buffer(hpm_renderer* r, int which, void* host_out, size_t bytes);
int hpm_write_buffer(hpm_renderer* r, int which, const void* host_in, size_t bytes);

/* Self-test: number of RNG-reachable arguments (1 - k*2^-23, k < 2^23) on which the tracker's branch-free logf differs from the
 * CUDA math library's logf (expected 0). */
int hpm_selftest_logf(uint64_t* mismatches_out);

/* Reference::Result (include/engine/graphics/Reference.hpp:17-29) -- what Reference::CompareNrc / CompareMc
 * (src/Reference.cpp:72-171; data/shader/ref/cmp1.comp, norm.comp, cmp2.comp) compute for a frame against a reference
 * frame, over the pixels whose reference alpha is not 0: mse = mean |cmp.rgb - ref.rgb|^2 / 3, the two image means,
 * own_var = mean |cmp.rgb - own_mean|^2 / 3.  Derived: bias = own_mean - ref_mean, rBias = bias / ref_mean,
 * rVar = own_var / ref_mean, CV = sqrt(own_var) / own_mean (Reference.cpp:10-28). */
typedef struct hpm_compare_result {
    float mse, ref_mean, own_mean, own_var;
    uint32_t valid_pixel_count;
} hpm_compare_result;
/* both images: device float[W*H][4] (HPM_BUF_OUTPUT layout).  Deterministic (fixed summation order). */
int hpm_compare_images(const float* d_ref_rgba, const float* d_cmp_rgba, uint32_t width, uint32_t height,
                       hpm_compare_result* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NRC_HPM_B200_H */
