// Host side of the Neural Radiance Cache: en::NeuralRadianceCache re-built on the sm_100a kernels of nrc_kernels.cuh,
// plus its C ABI (include/nrc_hpm_b200.h).  Mirrors reference src/NeuralRadianceCache.cu:11-178 (ctor / Init /
// InferAndTrain / Inference / Train / semaphores) and the slice of tiny-cuda-nn's Trainer it drives
// (trainer.h:50-87 initialisation, :163-211 training_step / loss; network_with_input_encoding.h:115-130 parameter order).
#include "nrc_host.h"
#include "nrc_kernels.cuh"
#include "nrc_wide_kernels.cuh"
#include "mini_json.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <string>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

namespace nrchpm {

std::atomic<uint64_t> g_launch_count{0};
static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }

// launch of a training-path kernel with the access-policy window that keeps [grad16 | w16] in the persisting part of L2
// `pdl`: programmatic dependent launch -- the kernel may be dispatched while the kernel before it on the stream drains; it orders
// itself behind that kernel with griddepcontrol.wait (pdl_wait) before it touches anything the kernel wrote
template <class K, class A>
void NrcCache::launch_hot(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const A& args, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (l2_window_bytes_) {
        attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[n].val.accessPolicyWindow.base_ptr = hot_.ptr;
        attr[n].val.accessPolicyWindow.num_bytes = l2_window_bytes_;
        attr[n].val.accessPolicyWindow.hitRatio = l2_hit_ratio_;
        attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        n++;
    }
    if (pdl && pdl_enabled_) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    if (n) { cfg.attrs = attr; cfg.numAttrs = n; }
    NRCHPM_CUDA(cudaLaunchKernelEx(&cfg, kernel, args));
}

// ------------------------------------------------------------------------------------------------ config
static int parse_pos(const mini_json::Value& v, NrcConfig& c) {
    const std::string t = v.string("otype", "");
    if (t == "HashGrid" || t == "Grid") {
        c.n_levels = (int)v.number("n_levels", 16);
        c.n_features = (int)v.number("n_features_per_level", 2);
        c.log2_hashmap_size = (int)v.number("log2_hashmap_size", 19);
        c.base_resolution = (int)v.number("base_resolution", 16);
        c.per_level_scale = (float)v.number("per_level_scale", 2.0);
        return POS_HASHGRID;
    }
    if (t == "Identity") return POS_IDENTITY;
    if (t == "TriangleWave") { c.n_freq_pos = (int)v.number("n_frequencies", 12); return POS_TRIANGLE; }
    if (t == "Frequency") { c.n_freq_pos = (int)v.number("n_frequencies", 12); return POS_FREQUENCY; }
    throw Error(NRCHPM_ERR_UNSUPPORTED, "position encoding '" + t + "' is not one of the reference presets (src/AppConfig.cpp:16-48)");
}
static int parse_dir(const mini_json::Value& v, NrcConfig& c) {
    const std::string t = v.string("otype", "");
    if (t == "OneBlob") { c.n_bins = (int)v.number("n_bins", 4); return DIR_ONEBLOB; }
    if (t == "Identity") return DIR_IDENTITY;
    if (t == "TriangleWave") { c.n_freq_dir = (int)v.number("n_frequencies", 4); return DIR_TRIANGLE; }
    throw Error(NRCHPM_ERR_UNSUPPORTED, "direction encoding '" + t + "' is not one of the reference presets (src/AppConfig.cpp:50-73)");
}

NrcConfig NrcConfig::from_json(const std::string& text) {
    NrcConfig c;
    mini_json::Value j;
    try { j = mini_json::parse(text); } catch (const std::exception& e) { throw Error(NRCHPM_ERR_INVALID, e.what()); }
    if (j.has("loss")) {
        const std::string t = j.at("loss").string("otype", "RelativeL2Luminance");
        if (t != "RelativeL2Luminance") throw Error(NRCHPM_ERR_UNSUPPORTED, "loss '" + t + "': only RelativeL2Luminance (the reference default, src/main.cu:433) is implemented");
    }
    if (j.has("optimizer")) {
        const auto& o = j.at("optimizer");
        const mini_json::Value* adam = &o;
        if (o.string("otype", "") == "EMA") {
            c.ema_decay = (float)o.number("decay", 0.99);
            if (o.has("nested")) adam = &o.at("nested");
        }
        const std::string t = adam->string("otype", "Adam");
        if (t != "Adam") throw Error(NRCHPM_ERR_UNSUPPORTED, "optimizer '" + t + "': only EMA(Adam) (src/NeuralRadianceCache.cu:20-27) is implemented");
        c.learning_rate = (float)adam->number("learning_rate", 1e-3);
        c.beta1 = (float)adam->number("beta1", 0.9); c.beta2 = (float)adam->number("beta2", 0.999);
        c.epsilon = (float)adam->number("epsilon", 1e-8); c.l2_reg = (float)adam->number("l2_reg", 1e-8);
    }
    NRCHPM_REQUIRE(j.has("encoding"), "config has no 'encoding'");
    const auto& enc = j.at("encoding");
    NRCHPM_REQUIRE(enc.string("otype", "") == "Composite" && enc.has("nested") && enc.at("nested").arr.size() == 2,
                   "encoding must be Composite{nested:[position, direction]} (src/AppConfig.cpp:75-80)");
    c.pos_enc = parse_pos(enc.at("nested").arr[0], c);
    c.dir_enc = parse_dir(enc.at("nested").arr[1], c);
    if (j.has("network")) {
        const auto& n = j.at("network");
        c.n_neurons = (int)n.number("n_neurons", 64);
        c.n_hidden_layers = (int)n.number("n_hidden_layers", 5);
        const std::string act = n.string("activation", "ReLU"), oact = n.string("output_activation", "None");
        if (act != "ReLU" || oact != "None") throw Error(NRCHPM_ERR_UNSUPPORTED, "only activation ReLU / output_activation None (src/NeuralRadianceCache.cu:30-36)");
    }
    c.infer_batch_size = (uint32_t)j.number("infer_batch_size", (double)(1u << 21));
    c.train_batch_size = (uint32_t)j.number("train_batch_size", (double)(1u << 14));
    c.train_batch_count = (uint32_t)j.number("train_batch_count", 4);
    if (j.has("compat")) c.oneblob_soa_bug = j.at("compat").boolean("oneblob_soa_bug", true);
    return c;
}

// ------------------------------------------------------------------------------------------------ derived layout
static float grid_scale(uint32_t level, float log2_pls, uint32_t base) { return exp2f((float)level * log2_pls) * (float)base - 1.0f; }   // common_device.h:700-706
static uint32_t grid_resolution(float scale) { return (uint32_t)ceilf(scale) + 1; }                                                       // common_device.h:708-710

void NrcCache::derive() {
    const NrcConfig& c = cfg_;
    NRCHPM_REQUIRE(c.n_neurons == 64 || c.n_neurons == 128, "n_neurons must be 64 or 128 (the widths tiny-cuda-nn's FullyFusedMLP is used with by the reference, src/AppConfig.cpp:169)");
    NRCHPM_REQUIRE(c.n_hidden_layers >= 1 && c.n_hidden_layers <= 8, "n_hidden_layers must be in [1, 8]");
    NRCHPM_REQUIRE(c.n_features == 2, "HashGrid n_features_per_level must be 2");
    NRCHPM_REQUIRE(c.n_levels >= 1 && c.n_levels <= kMaxLevels, "HashGrid n_levels must be in [1, 16]");
    NRCHPM_REQUIRE(c.n_freq_pos >= 1 && c.n_freq_pos <= 12 && c.n_freq_dir >= 1 && c.n_freq_dir <= 8 && c.n_bins >= 1 && c.n_bins <= 8, "encoding sizes out of range");
    EncParams& e = enc_;
    std::memset(&e, 0, sizeof(e));
    e.pos_enc = c.pos_enc; e.dir_enc = c.dir_enc;
    e.n_levels = c.n_levels; e.n_freq_pos = c.n_freq_pos; e.n_freq_dir = c.n_freq_dir; e.n_bins = c.n_bins;
    const int pw[4] = {c.n_levels * 2, 3, 3 * c.n_freq_pos, 3 * c.n_freq_pos * 2};
    const int dw[3] = {2 * c.n_bins, 2, 2 * c.n_freq_dir};
    e.pos_w = pw[c.pos_enc]; e.dir_w = dw[c.dir_enc];
    e.dir_off = e.pos_w;
    e.in_w = ((e.pos_w + e.dir_w + 15) / 16) * 16;                      // set_alignment(16), network_with_input_encoding.h:47
    NRCHPM_REQUIRE(e.in_w <= 80, "encoded width > 80 not supported");
    e.oneblob_soa = (c.pos_enc == POS_HASHGRID) ? 1 : 0;                // composite.h:400-403
    e.soa_bug = (e.oneblob_soa && c.dir_enc == DIR_ONEBLOB && c.oneblob_soa_bug) ? 1 : 0;
    n_grid_ = 0;
    if (c.pos_enc == POS_HASHGRID) {                                    // grid.h:699-724
        uint32_t offset = 0;
        e.all_pow2 = 1;
        const float log2_pls = std::log2(c.per_level_scale);
        for (int i = 0; i < c.n_levels; i++) {
            const float scale = grid_scale(i, log2_pls, c.base_resolution);
            const uint32_t res = grid_resolution(scale);
            const uint32_t max_params = 0xFFFFFFFFu / 2;
            uint32_t p = std::pow((float)res, 3) > (float)max_params ? max_params : res * res * res;
            p = ((p + 7) / 8) * 8;
            p = std::min(p, 1u << c.log2_hashmap_size);
            e.level_scale[i] = scale;
            e.level_hsize[i] = p;
            if (p & (p - 1)) e.all_pow2 = 0;
            e.level_offset[i] = offset;
            // grid_index (common_device.h:668-690) with its uint32 stride arithmetic, including the wrap-around for res >= 2^16
            uint32_t stride = 1, s[3] = {0, 0, 0};
            for (uint32_t dim = 0; dim < 3 && stride <= p; ++dim) { s[dim] = stride; stride *= res; }
            e.level_s0[i] = s[0]; e.level_s1[i] = s[1]; e.level_s2[i] = s[2];
            e.level_hash[i] = p < stride ? 1u : 0u;
            {   // kind of index arithmetic (nrc_kernels.cuh: hashgrid_pipelined); "generic" covers wrapped strides and odd table sizes
                const bool pow2 = (p & (p - 1)) == 0;
                const bool dup = !e.level_hash[i] && pow2 && ((s[2] & (p - 1)) == 0 || (s[1] & (p - 1)) == 0);
                e.level_kind[i] = e.level_hash[i] ? 2u : (dup || !pow2) ? 0u : 1u;
            }
            offset += p;
        }
        n_grid_ = (size_t)offset * 2;
    }
    const int W = c.n_neurons;
    n_mlp_ = (size_t)W * e.in_w + (size_t)(c.n_hidden_layers - 1) * W * W + (size_t)kOutPad * W;
    n_params_ = n_mlp_ + n_grid_;
}

// pcg32 as vendored by tiny-cuda-nn (dependencies/pcg32/pcg32.h)
namespace {
struct Pcg32 {
    uint64_t state, inc;
    explicit Pcg32(uint64_t initstate, uint64_t initseq = 1u) { state = 0u; inc = (initseq << 1u) | 1u; next_uint(); state += initstate; next_uint(); }
    uint32_t next_uint() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    float next_float() { uint32_t u = (next_uint() >> 9) | 0x3f800000u; float f; std::memcpy(&f, &u, 4); return f - 1.0f; }
};
}  // namespace

// Trainer::initialize_params (trainer.h:68-87): MLP xavier-uniform matrix by matrix on the host (gpu_matrix.h:284-299,
// fully_fused_mlp.cu:866-891), grid U(-1e-4, 1e-4) with the device generator's element order (random.h:40-66).
void NrcCache::init_params(uint64_t seed) {
    std::vector<float> master(n_params_);
    std::seed_seq seq{(uint32_t)seed};
    std::vector<uint32_t> seeds(2);
    seq.generate(seeds.begin(), seeds.end());
    Pcg32 rng{seeds.front()};
    const int W = cfg_.n_neurons, H = cfg_.n_hidden_layers;
    size_t off = 0;
    auto fill = [&](int rows, int cols) {
        const float scale = std::sqrt(6.0f / (float)(cols + rows));
        for (size_t i = 0; i < (size_t)rows * cols; i++) master[off + i] = rng.next_float() * 2.0f * scale - scale;
        off += (size_t)rows * cols;
    };
    fill(W, enc_.in_w);
    for (int i = 0; i < H - 1; i++) fill(W, W);
    fill(kOutPad, W);
    if (n_grid_) {
        const size_t n = n_grid_;
        const size_t n_threads = 128 * ((((n + 3) / 4) + 127) / 128);
        float* g = master.data() + n_mlp_;
        Pcg32 base = rng;
        for (size_t i = 0; i < n_threads; i++) {
            Pcg32 r = base;
            for (int j = 0; j < 4; j++) {
                const size_t idx = i + n_threads * j;
                const float v = r.next_float();
                if (idx < n) g[idx] = fmaf(v, 2e-4f, -1e-4f);
            }
            base.next_uint(); base.next_uint(); base.next_uint(); base.next_uint();
        }
    }
    set_params_fp32(master.data());
    NRCHPM_CUDA(cudaMemset(ema16_.ptr, 0, ema16_.bytes()));             // ema.h:93-94
    NRCHPM_CUDA(cudaMemset(m1_.ptr, 0, m1_.bytes()));
    NRCHPM_CUDA(cudaMemset(m2_.ptr, 0, m2_.bytes()));
    NRCHPM_CUDA(cudaMemset(steps_.ptr, 0, steps_.bytes()));
    if (n_grid_) {      // moments and step counters of the encoding: zero everything but the master weights just written
        DeviceBuffer<float> zeros;
        zeros.allocate(n_grid_);
        zeros.zero();
        for (int field = 1; field <= 3; field++) scatter_grid_field(field, zeros.ptr);
        NRCHPM_CUDA(cudaDeviceSynchronize());
    }
    NRCHPM_CUDA(cudaMemset(hot_.ptr, 0, grad16_.bytes()));
    current_step_ = 0;
}

NrcCache::NrcCache(const NrcConfig& cfg, uint64_t seed) : cfg_(cfg) {
    derive();
    int dev = 0;
    NRCHPM_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    NRCHPM_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) throw Error(NRCHPM_ERR_CUDA, std::string("this library contains sm_100a code only; device is ") + prop.name);
    sm_count_ = prop.multiProcessorCount;
    {
        const size_t n_pad = (n_params_ + 63) / 64 * 64;
        hot_.allocate(2 * n_pad);
        grad16_.ptr = hot_.ptr; grad16_.count = n_params_;               // first: the cudaIpc handle of the gradient is the allocation's
        w16_.ptr = hot_.ptr + n_pad; w16_.count = n_params_;
        ema16_.allocate(n_params_);
        const char* v = std::getenv("NRCHPM_L2_PERSIST");
        if (n_grid_ && !(v && std::atoi(v) == 0) && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            const size_t want = std::min<size_t>(hot_.bytes(), (size_t)prop.accessPolicyMaxWindowSize);
            const size_t aside = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, want);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, aside) == cudaSuccess) {
                l2_window_bytes_ = want;
                l2_hit_ratio_ = (float)std::min(1.0, (double)aside / (double)want);
            } else cudaGetLastError();
        }
    }
    master_.allocate(n_mlp_); m1_.allocate(n_mlp_); m2_.allocate(n_mlp_); steps_.allocate(n_mlp_);      // network weights: SoA
    grid_state_.allocate(n_grid_ / 2);                                                                    // encoding: one record per entry
    loss_dev_.allocate(1);
    NRCHPM_CUDA(cudaMemset(loss_dev_.ptr, 0, sizeof(float)));
    init_params(seed);
    setup_kernels();
}

NrcCache::~NrcCache() {
    for (auto e : pipe_events_) cudaEventDestroy(e);
    if (loss_pinned_) cudaFreeHost(loss_pinned_);
    if (ov_inf_stream_) { cudaStreamDestroy(ov_inf_stream_); cudaStreamDestroy(ov_tr_stream_); for (auto e : ov_ev_) cudaEventDestroy(e); }
    if (ema_stream_) {
        cudaStreamSynchronize(ema_stream_);
        cudaStreamDestroy(ema_stream_);
        cudaEventDestroy(adam_done_); cudaEventDestroy(ema_done_);
    }
    if (copy_in_stream_) { cudaStreamDestroy(copy_in_stream_); cudaStreamDestroy(copy_out_stream_); cudaStreamDestroy(compute_stream_); cudaStreamDestroy(train_stream_); }
}

void NrcCache::scatter_grid_field(int field, const float* d_src) {
    const uint64_t n = n_grid_ / 2;
    nrc_grid_state_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, nullptr>>>(grid_state_.ptr, n, field, d_src);
    check_launch("nrc_grid_state_scatter_kernel");
}

void NrcCache::set_params_fp32(const float* host_master) {
    NRCHPM_CUDA(cudaDeviceSynchronize());
    NRCHPM_CUDA(cudaMemcpy(master_.ptr, host_master, n_mlp_ * sizeof(float), cudaMemcpyHostToDevice));
    if (n_grid_) {
        DeviceBuffer<float> tmp;
        tmp.allocate(n_grid_);
        NRCHPM_CUDA(cudaMemcpy(tmp.ptr, host_master + n_mlp_, n_grid_ * sizeof(float), cudaMemcpyHostToDevice));
        scatter_grid_field(0, tmp.ptr);
        NRCHPM_CUDA(cudaDeviceSynchronize());
    }
    std::vector<__half> h(n_params_);
    for (size_t i = 0; i < n_params_; i++) h[i] = __float2half_rn(host_master[i]);
    NRCHPM_CUDA(cudaMemcpy(w16_.ptr, h.data(), n_params_ * sizeof(__half), cudaMemcpyHostToDevice));
}

void NrcCache::set_ema(const float* host_ema) {
    NRCHPM_CUDA(cudaDeviceSynchronize());          // an EMA pass may still be running on its side stream
    std::vector<__half> h(n_params_);
    for (size_t i = 0; i < n_params_; i++) h[i] = __float2half_rn(host_ema[i]);
    NRCHPM_CUDA(cudaMemcpy(ema16_.ptr, h.data(), n_params_ * sizeof(__half), cudaMemcpyHostToDevice));
}

void NrcCache::get_params(int which, float* out) {
    if (which == 3 && grads_pending_) {       // before the optimizer ran, the MLP gradient only exists as per-chunk partials
        const float* src = dw_source_ ? dw_source_ : dw_partials_.ptr;
        nrc_partials_to_half_kernel<<<(unsigned)((n_mlp_ + 255) / 256), 256, 0, train_stream_last_>>>(src, dw_source_ ? 1u : dw_chunks_, (uint32_t)n_mlp_, grad16_.ptr);
        check_launch("nrc_partials_to_half_kernel");
    }
    NRCHPM_CUDA(cudaDeviceSynchronize());
    if (which == 0 || which == 4 || which == 5 || which == 6) {
        // network part: SoA arrays; encoding part: gathered out of the per-entry records into tcnn's parameter order
        if (which == 6) {
            std::vector<uint32_t> h(n_mlp_);
            NRCHPM_CUDA(cudaMemcpy(h.data(), steps_.ptr, n_mlp_ * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n_mlp_; i++) out[i] = (float)h[i];
        } else {
            const float* src = which == 0 ? master_.ptr : which == 4 ? m1_.ptr : m2_.ptr;
            NRCHPM_CUDA(cudaMemcpy(out, src, n_mlp_ * sizeof(float), cudaMemcpyDeviceToHost));
        }
        if (n_grid_) {
            DeviceBuffer<float> tmp;
            tmp.allocate(n_grid_);
            const uint64_t n = n_grid_ / 2;
            const int field = which == 0 ? 0 : which == 4 ? 1 : which == 5 ? 2 : 3;
            nrc_grid_state_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, nullptr>>>(grid_state_.ptr, n, field, tmp.ptr);
            check_launch("nrc_grid_state_gather_kernel");
            NRCHPM_CUDA(cudaMemcpy(out + n_mlp_, tmp.ptr, n_grid_ * sizeof(float), cudaMemcpyDeviceToHost));
        }
    } else if (which >= 1 && which <= 3) {
        std::vector<__half> h(n_params_);
        const __half* src = which == 1 ? w16_.ptr : which == 2 ? ema16_.ptr : grad16_.ptr;
        NRCHPM_CUDA(cudaMemcpy(h.data(), src, n_params_ * sizeof(__half), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n_params_; i++) out[i] = __half2float(h[i]);
    } else throw Error(NRCHPM_ERR_INVALID, "nrc_get_params: which must be 0..6");
}

// ------------------------------------------------------------------------------------------------ kernel dispatch
#define NRC_DISPATCH_INW(inw, ...)                                                         \
    switch (inw) {                                                                         \
        case 16: { constexpr int IN_W = 16; __VA_ARGS__; } break;                          \
        case 32: { constexpr int IN_W = 32; __VA_ARGS__; } break;                          \
        case 48: { constexpr int IN_W = 48; __VA_ARGS__; } break;                          \
        case 64: { constexpr int IN_W = 64; __VA_ARGS__; } break;                          \
        case 80: { constexpr int IN_W = 80; __VA_ARGS__; } break;                          \
        default: throw Error(NRCHPM_ERR_UNSUPPORTED, "unsupported encoded width");         \
    }

void NrcCache::setup_kernels() {
    const int H = cfg_.n_hidden_layers;
    if (wide()) {
        // 128 neurons (nrc_wide_kernels.cuh): the resident weight image bounds the depth; two tiles per CTA if their input tiles still fit
        NRC_DISPATCH_INW(enc_.in_w, {
            NRCHPM_REQUIRE(wide_fwd_smem_bytes<IN_W>(H, 1) <= 227 * 1024, "128-neuron network: at most 6 hidden layers fit the shared-memory weight image");
            wide_wgs_ = wide_fwd_smem_bytes<IN_W>(H, 2) <= 227 * 1024 ? 2 : 1;
            NRCHPM_CUDA(cudaFuncSetAttribute(nrc_wide_forward_kernel<IN_W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_fwd_smem_bytes<IN_W>(H, wide_wgs_)));
            NRCHPM_CUDA(cudaFuncSetAttribute(nrc_wide_forward_kernel<IN_W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_fwd_smem_bytes<IN_W>(H, wide_wgs_)));
            NRCHPM_CUDA(cudaFuncSetAttribute(nrc_wide_backward_kernel<IN_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_bwd_smem_bytes<IN_W>(H)));
            NRCHPM_CUDA(cudaFuncSetAttribute(nrc_wide_dw_kernel<IN_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWideDwSmemBytes));
            wide_ws_slots_ = wide_ws_slots<IN_W>(H);
            if (wide_ws_slots_ >= 2)
                NRCHPM_CUDA(cudaFuncSetAttribute(nrc_wide_infer_ws_kernel<IN_W, NRC_WIDE_WS_NP, NRC_WIDE_WS_NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_ws_smem_bytes<IN_W>(H, wide_ws_slots_)));
        });
    }
    NRC_DISPATCH_INW(enc_.in_w, {
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_forward_kernel<IN_W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(fwd_smem_bytes<IN_W>(H, kInferWgs) + infer_smem_level_bytes())));
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_forward_kernel<IN_W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem_bytes<IN_W>(H)));
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_backward_kernel<IN_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes<IN_W>(H)));
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_dw_kernel<IN_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDwSmemBytes));
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_forward2_kernel<IN_W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd2_smem_bytes<IN_W>(H, 3)));
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_forward2_kernel<IN_W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd2_smem_bytes<IN_W>(H, 1)));
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_backward2_kernel<IN_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd2_smem_bytes<IN_W>(H)));
        NRCHPM_CUDA(cudaFuncSetAttribute(nrc_infer_ws_kernel<IN_W, NRC_WS_NP, NRC_WS_NC, ws_slots(IN_W)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)infer_ws_smem_bytes<IN_W>(H, ws_slots(IN_W))));
        if (const char* v = std::getenv("NRCHPM_WS_CARVEOUT"))      // experiment: shared-memory carve-out (percent) of the SMs that run the inference kernel
            NRCHPM_CUDA(cudaFuncSetAttribute(nrc_infer_ws_kernel<IN_W, NRC_WS_NP, NRC_WS_NC, ws_slots(IN_W)>, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(v)));
        if (fused_training_fits()) {
            NRCHPM_CUDA(cudaFuncSetAttribute(nrc_train_fused_kernel<IN_W, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)train_smem_bytes<IN_W>(H)));
            NRCHPM_CUDA(cudaFuncSetAttribute(nrc_train_fused_kernel<IN_W, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)train_smem_bytes<IN_W>(H)));
        }
    });
    NRCHPM_CUDA(cudaFuncSetAttribute(nrc_peer_adam_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)peer_adam_smem_bytes<2>()));
    NRCHPM_CUDA(cudaFuncSetAttribute(nrc_peer_adam_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)peer_adam_smem_bytes<4>()));
    NRCHPM_CUDA(cudaFuncSetAttribute(nrc_peer_adam_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)peer_adam_smem_bytes<8>()));
    // tuning knobs (experiments only; the defaults are the measured best)
    if (const char* v = std::getenv("NRCHPM_INFER_GROUPS")) infer_groups_ = std::max(0, std::min(3, std::atoi(v)));
    if (const char* v = std::getenv("NRCHPM_PDL")) pdl_enabled_ = std::atoi(v) != 0;                   // programmatic dependent launch of the training chain
    if (const char* v = std::getenv("NRCHPM_TRAIN_GROUPS")) train_groups_ = std::max(0, std::min(1, std::atoi(v)));
    if (const char* v = std::getenv("NRCHPM_INFER_WS")) infer_ws_ = std::atoi(v) != 0 ? 1 : 0;   // 0: tile-per-warpgroup kernel
    if (const char* v = std::getenv("NRCHPM_OVERLAP")) overlap_schedule_ = std::atoi(v) != 0;          // data-parallel replicas: 0 = serial Inference() -> Train()
    if (const char* v = std::getenv("NRCHPM_OVERLAP_SMS")) overlap_infer_sms_ = (uint32_t)std::max(1, std::atoi(v));
    if (const char* v = std::getenv("NRCHPM_OVERLAP_HEAD")) overlap_head_ = std::max(0.0, std::min(1.0, std::atof(v)));
    if (const char* v = std::getenv("NRCHPM_PEER_FUSED")) peer_fused_ = std::atoi(v) != 0;             // 0: gather / Adam / publish as three kernels
    if (const char* v = std::getenv("NRCHPM_PEER_CTAS")) peer_ctas_ = (uint32_t)std::max(1, std::atoi(v));
    if (const char* v = std::getenv("NRCHPM_TRAIN_FUSED")) train_fused_ = std::atoi(v) != 0;          // 0: the three-kernel path
    if (const char* v = std::getenv("NRCHPM_TRAIN_TPR")) train_tpr_ = std::atoi(v) == 4 ? 4 : 2;      // threads per record of the fused kernel
    if (const char* v = std::getenv("NRCHPM_TRAIN_PROF")) if (std::atoi(v)) { train_prof_.allocate((size_t)sm_count_ * 16); train_prof_.zero(); timeline_.allocate(2 * kTimelineSlots); reset_timeline(); }   // development aid
}

// the fused training kernel keeps one fp32 accumulator region per weight matrix in tensor memory and every hidden layer's
// activations in shared memory: networks of up to six hidden layers (every configuration of the thesis) fit, deeper ones take the
// three-kernel path
bool NrcCache::fused_training_fits() const {
    const int H = cfg_.n_hidden_layers;
    size_t smem = 0;
    NRC_DISPATCH_INW(enc_.in_w, { smem = train_smem_bytes<IN_W>(H); });
    return !wide() && train_tmem_cols(enc_.in_w, H) <= 512 && smem <= 227 * 1024;
}

// bytes of the coarse hash-grid levels the inference kernel stages in shared memory (0: none)
size_t NrcCache::infer_smem_level_bytes() const {
    if (NRC_SMEM_LEVELS <= 0 || enc_.pos_enc != POS_HASHGRID || enc_.n_levels <= NRC_SMEM_LEVELS) return 0;
    const size_t bytes = (size_t)enc_.level_offset[NRC_SMEM_LEVELS] * 4;
    return (bytes % 16 == 0 && bytes <= 32768) ? bytes : 0;
}

// Parameter snapshot for inference that overlaps training (multi-GPU frames: the cache is evaluated while the gradient
// all-reduce of a training step is in flight; Inference() must still see the parameters of the previous frame).
void NrcCache::snapshot_params(bool use_ema, cudaStream_t s) {
    infer_snapshot_.ensure(n_params_);
    if (use_ema) wait_ema(s);
    NRCHPM_CUDA(cudaMemcpyAsync(infer_snapshot_.ptr, use_ema ? ema16_.ptr : w16_.ptr, n_params_ * sizeof(__half), cudaMemcpyDeviceToDevice, s));
    snapshot_valid_ = true;
}
// param_set: 0 working weights, 1 EMA weights (the reference's Inference()), 2 the last snapshot
void NrcCache::inference_set(int param_set, const float* d_in, float* d_out, uint32_t n, const uint32_t* d_indices, const uint32_t* d_count, cudaStream_t s) {
    if (param_set != 2) { inference(d_in, d_out, n, param_set != 0, d_indices, d_count, s); return; }
    NRCHPM_REQUIRE(snapshot_valid_, "inference from the snapshot without nrc_snapshot_params");
    inference_with(infer_snapshot_.ptr, d_in, d_out, n, d_indices, d_count, s, 0);
}

void NrcCache::encode(const float* d_in, uint32_t n, bool use_ema, void* d_out_half, cudaStream_t s) {
    if (n == 0) return;
    if (use_ema) wait_ema(s);
    nrc_encode_kernel<<<(n + 127) / 128, 128, 0, s>>>(enc_, use_ema ? ema16_.ptr : w16_.ptr, (uint32_t)n_mlp_, d_in, n, (__half*)d_out_half);
    check_launch("nrc_encode_kernel");
}

void NrcCache::inference(const float* d_in, float* d_out, uint32_t n, bool use_ema, const uint32_t* d_indices, const uint32_t* d_count, cudaStream_t s) {
    if (n == 0) return;
    if (use_ema) wait_ema(s);
    inference_with(use_ema ? ema16_.ptr : w16_.ptr, d_in, d_out, n, d_indices, d_count, s, infer_max_ctas_);
}

// the inference launch on an explicit fp16 parameter vector [network | encoding]; max_ctas caps the persistent grid (0: whole GPU)
void NrcCache::inference_with(const __half* params, const float* d_in, float* d_out, uint32_t n, const uint32_t* d_indices, const uint32_t* d_count, cudaStream_t s,
                              uint32_t max_ctas) {
    if (n == 0) return;
    FwdArgs a{};
    a.enc = enc_; a.params = params; a.n_mlp = (uint32_t)n_mlp_; a.n_hidden = cfg_.n_hidden_layers;
    a.in = d_in; a.indices = d_indices; a.d_count = d_count; a.n = n; a.out = d_out;
    a.tl = timeline_slot();
    const uint32_t tiles = (n + kTile - 1) / kTile;
    uint32_t grid, threads;
    if (wide() && infer_ws_ > 0 && enc_.pos_enc == POS_HASHGRID && wide_ws_slots_ >= 2 && tiles >= (uint32_t)sm_count_) {
        // 128 neurons, hash grid: warp-specialised persistent kernel (producer warpgroups gather, consumer warpgroups run the MLP)
        grid = (uint32_t)sm_count_;
        if (max_ctas) grid = std::min(grid, max_ctas);
        a.ring_slots = (uint32_t)wide_ws_slots_;
        NRC_DISPATCH_INW(enc_.in_w, { launch_hot(nrc_wide_infer_ws_kernel<IN_W, NRC_WIDE_WS_NP, NRC_WIDE_WS_NC>, grid, (NRC_WIDE_WS_NP + NRC_WIDE_WS_NC) * 128,
                                                 wide_ws_smem_bytes<IN_W>(cfg_.n_hidden_layers, wide_ws_slots_), s, a); });
        check_launch("nrc_wide_infer_ws_kernel");
        return;
    }
    if (wide()) {
        // 128 neurons: persistent grid, one CTA per SM (176 KB weight image), `wide_wgs_` tiles in flight per CTA
        const uint32_t wgs = tiles >= (uint32_t)sm_count_ * 2 ? (uint32_t)wide_wgs_ : 1u;
        grid = std::min<uint32_t>((tiles + wgs - 1) / wgs, (uint32_t)sm_count_);
        if (max_ctas) grid = std::min(grid, max_ctas);
        NRC_DISPATCH_INW(enc_.in_w, { launch_hot(nrc_wide_forward_kernel<IN_W, false>, grid, wgs * 128, wide_fwd_smem_bytes<IN_W>(cfg_.n_hidden_layers, (int)wgs), s, a); });
        check_launch("nrc_wide_forward_kernel<infer>");
        return;
    }
    if (infer_groups_ > 0) {
        const uint32_t g = (uint32_t)infer_groups_;
        grid = std::min<uint32_t>((tiles + g - 1) / g, (uint32_t)sm_count_);
        NRC_DISPATCH_INW(enc_.in_w, {
            nrc_forward2_kernel<IN_W, false><<<grid, g * kGroupThreads, fwd2_smem_bytes<IN_W>(cfg_.n_hidden_layers, (int)g), s>>>(a);
        });
        check_launch("nrc_forward2_kernel<infer>");
        return;
    }
    if (infer_ws_ > 0 && enc_.pos_enc == POS_HASHGRID && tiles >= (uint32_t)sm_count_) {
        // warp-specialised persistent kernel: one CTA per SM, producer warpgroups encode, consumer warpgroups run the MLP
        grid = (uint32_t)sm_count_;
        if (max_ctas) grid = std::min(grid, max_ctas);
        // (same access-policy window as the training launches: kernels with different windows do not share the GPU)
        NRC_DISPATCH_INW(enc_.in_w, { launch_hot(nrc_infer_ws_kernel<IN_W, NRC_WS_NP, NRC_WS_NC, ws_slots(IN_W)>, grid, (NRC_WS_NP + NRC_WS_NC) * 128, infer_ws_smem_bytes<IN_W>(cfg_.n_hidden_layers, ws_slots(IN_W)), s, a); });
        check_launch("nrc_infer_ws_kernel");
        return;
    }
    // persistent grid: kInferWgs warpgroups (tiles in flight) per CTA, kInferCtas CTAs per SM; small batches use fewer warpgroups
    const uint32_t wgs = (uint32_t)std::max(1, std::min<int>(kInferWgs, (int)((tiles + sm_count_ * kInferCtas - 1) / (sm_count_ * kInferCtas))));
    threads = wgs * 128;
    grid = std::min<uint32_t>((tiles + wgs - 1) / wgs, (uint32_t)sm_count_ * kInferCtas);
    if (max_ctas) grid = std::min(grid, max_ctas);          // leave room for a co-running kernel (NCCL all-reduce)
    const size_t lvl_bytes = infer_smem_level_bytes();
    if (lvl_bytes) { a.smem_levels = NRC_SMEM_LEVELS; a.smem_level_entries = (uint32_t)(lvl_bytes / 4); }
    NRC_DISPATCH_INW(enc_.in_w, {
        launch_hot(nrc_forward_kernel<IN_W, false>, grid, threads, fwd_smem_bytes<IN_W>(cfg_.n_hidden_layers, (int)wgs) + lvl_bytes, s, a);
    });
    check_launch("nrc_forward_kernel<infer>");
}

// Persistent grid: two warpgroups (tiles) per CTA, at most two CTAs per SM.  (One warpgroup per CTA on twice as many SMs was
// measured slower for the 128-tile training batch: 81 vs 62 us backward, the per-CTA weight prologue dominates.)
void NrcCache::launch_shape(uint32_t tiles, uint32_t& grid, uint32_t& threads) const {
    threads = kFwdThreads;
    grid = std::min<uint32_t>((tiles + 1) / 2, (uint32_t)sm_count_ * 2);
}

void NrcCache::ensure_train_scratch(uint32_t B) {
    if (B <= scratch_batch_) return;
    const int H = cfg_.n_hidden_layers;
    if (!(train_fused_ && fused_training_fits())) {       // the fused kernel keeps these in shared / tensor memory
        x16_.allocate((size_t)B * enc_.in_w);
        acts_.allocate((size_t)H * B * cfg_.n_neurons);
        dacts_.allocate((size_t)H * B * cfg_.n_neurons);
        dx16_.allocate((size_t)B * enc_.in_w);
    }
    out16_.allocate((size_t)B * kOutPad);
    dout16_.allocate((size_t)B * kOutPad);
    if (!train_done_.ptr) { train_done_.allocate(1); train_done_.zero(); }
    loss_partials_.allocate(B / kTile);
    dw_partials_.allocate((size_t)kMaxDwChunks * n_mlp_);
    scratch_batch_ = B;
}

void NrcCache::training_step(const float* d_in, const float* d_target, uint32_t B, bool run_optimizer, cudaStream_t s) {
    NRCHPM_REQUIRE(B > 0 && B % kTile == 0, "training batch must be a positive multiple of 128 (tcnn: 256, common.h:235)");
    ensure_train_scratch(B);
    const int H = cfg_.n_hidden_layers;
    if (grid_grad_dirty_ && n_grid_) NRCHPM_CUDA(cudaMemsetAsync(grad16_.ptr + n_mlp_, 0, n_grid_ * sizeof(__half), s));   // grid.h:857-860
    const uint32_t tiles = B / kTile;
    if (train_fused_ && fused_training_fits()) {
        // one launch: forward, loss, backward, weight gradients (per-CTA fp32 partials) and the hash-grid gradient scatter
        TrainArgs a{};
        a.enc = enc_; a.params = w16_.ptr; a.n_mlp = (uint32_t)n_mlp_; a.n_hidden = H;
        a.in = d_in; a.target = d_target; a.n = B; a.loss_scale = cfg_.loss_scale;
        a.grid_grad = n_grid_ ? grad16_.ptr + n_mlp_ : nullptr;
        a.dw_partials = dw_partials_.ptr; a.loss_partials = loss_partials_.ptr; a.loss_out = loss_dev_.ptr; a.done_counter = train_done_.ptr;
        a.out16 = out16_.ptr; a.dout16 = dout16_.ptr;
        a.grid_state = n_grid_ ? grid_state_.ptr : nullptr;
        a.prof = train_prof_.ptr;
        a.tl = timeline_slot();
        const uint32_t grid = std::min<uint32_t>(tiles, (uint32_t)sm_count_);
        if (train_tpr_ == 4) { NRC_DISPATCH_INW(enc_.in_w, { launch_hot(nrc_train_fused_kernel<IN_W, 4>, grid, 512, train_smem_bytes<IN_W>(H), s, a, true); }); }
        else { NRC_DISPATCH_INW(enc_.in_w, { launch_hot(nrc_train_fused_kernel<IN_W, 2>, grid, 256, train_smem_bytes<IN_W>(H), s, a, true); }); }
        check_launch("nrc_train_fused_kernel");
        dw_chunks_ = grid;
        grid_grad_dirty_ = n_grid_ != 0;
    } else {
        training_step_three_kernels(d_in, d_target, B, s);
    }
    last_batch_ = B;
    dw_source_ = nullptr;
    loss_valid_ = false;
    grads_pending_ = true;
    train_stream_last_ = s;
    if (run_optimizer) optimizer_step(s);
}

// networks that do not fit the fused kernel (more than six hidden layers): forward / backward / weight-gradient kernels with the
// activations in HBM between them
void NrcCache::training_step_three_kernels(const float* d_in, const float* d_target, uint32_t B, cudaStream_t s) {
    const int H = cfg_.n_hidden_layers;
    const uint32_t tiles = B / kTile;
    uint32_t grid, threads;
    launch_shape(tiles, grid, threads);
    if (wide()) {
        // 128 neurons: one CTA per SM; small batches spread one tile per CTA over the SMs, large ones keep two tiles in flight per CTA
        const uint32_t wgs = tiles >= (uint32_t)sm_count_ * 2 ? (uint32_t)wide_wgs_ : 1u;
        const uint32_t wgrid = std::min<uint32_t>((tiles + wgs - 1) / wgs, (uint32_t)sm_count_);
        FwdArgs f{};
        f.enc = enc_; f.params = w16_.ptr; f.n_mlp = (uint32_t)n_mlp_; f.n_hidden = H;
        f.in = d_in; f.n = B; f.target = d_target;
        f.x16 = x16_.ptr; f.acts = acts_.ptr; f.out16 = out16_.ptr; f.dout16 = dout16_.ptr; f.loss_partials = loss_partials_.ptr;
        f.loss_scale = cfg_.loss_scale;
        NRC_DISPATCH_INW(enc_.in_w, { launch_hot(nrc_wide_forward_kernel<IN_W, true>, wgrid, wgs * 128, wide_fwd_smem_bytes<IN_W>(H, (int)wgs), s, f); });
        check_launch("nrc_wide_forward_kernel<train>");
        BwdArgs b{};
        b.enc = enc_; b.params = w16_.ptr; b.n_mlp = (uint32_t)n_mlp_; b.n_hidden = H; b.n = B;
        b.in = d_in; b.acts = acts_.ptr; b.dout16 = dout16_.ptr; b.dacts = dacts_.ptr;
        b.need_dx = n_grid_ ? 1 : 0;
        b.dx16 = (n_grid_ && keep_dx_) ? dx16_.ptr : nullptr;
        b.grid_grad = n_grid_ ? grad16_.ptr + n_mlp_ : nullptr;
        b.loss_partials = loss_partials_.ptr; b.loss_out = loss_dev_.ptr; b.n_loss_partials = tiles;
        NRC_DISPATCH_INW(enc_.in_w, { launch_hot(nrc_wide_backward_kernel<IN_W>, wgrid, wgs * 128, wide_bwd_smem_bytes<IN_W>(H), s, b); });
        check_launch("nrc_wide_backward_kernel");
        grid_grad_dirty_ = n_grid_ != 0;
        DwArgs d{};
        d.n_hidden = H; d.n = B; d.n_mlp = (uint32_t)n_mlp_;
        // chunks x (H + 1) CTAs, one per SM (128 KB of operand stages): ~3 waves at the reference's 2^14 batch
        const uint32_t want_chunks = tiles <= 1024 ? std::min<uint32_t>(tiles, 64u) : std::min<uint32_t>(kMaxDwChunks, tiles / 16);
        const uint32_t tiles_per_chunk = (tiles + want_chunks - 1) / want_chunks;
        d.kc = tiles_per_chunk * kTile;
        dw_chunks_ = (B + d.kc - 1) / d.kc;
        d.x16 = x16_.ptr; d.acts = acts_.ptr; d.dacts = dacts_.ptr; d.dout16 = dout16_.ptr; d.partials = dw_partials_.ptr;
        NRC_DISPATCH_INW(enc_.in_w, { nrc_wide_dw_kernel<IN_W><<<dim3(dw_chunks_, H + 1), 128, kWideDwSmemBytes, s>>>(d); });
        check_launch("nrc_wide_dw_kernel");
        return;
    }
    {
        FwdArgs a{};
        a.enc = enc_; a.params = w16_.ptr; a.n_mlp = (uint32_t)n_mlp_; a.n_hidden = H;
        a.in = d_in; a.n = B; a.target = d_target;
        a.x16 = x16_.ptr; a.acts = acts_.ptr; a.out16 = out16_.ptr; a.dout16 = dout16_.ptr; a.loss_partials = loss_partials_.ptr;
        a.loss_scale = cfg_.loss_scale;
        if (train_groups_ > 0) {
            const uint32_t g = (uint32_t)train_groups_, grid2 = std::min<uint32_t>((tiles + g - 1) / g, (uint32_t)sm_count_ * 2);
            NRC_DISPATCH_INW(enc_.in_w, { nrc_forward2_kernel<IN_W, true><<<grid2, g * kGroupThreads, fwd2_smem_bytes<IN_W>(H, (int)g), s>>>(a); });
        } else {
            NRC_DISPATCH_INW(enc_.in_w, { nrc_forward_kernel<IN_W, true><<<grid, threads, fwd_smem_bytes<IN_W>(H), s>>>(a); });
        }
        check_launch("nrc_forward_kernel<train>");
    }
    {
        BwdArgs a{};
        a.enc = enc_; a.params = w16_.ptr; a.n_mlp = (uint32_t)n_mlp_; a.n_hidden = H; a.n = B;
        a.in = d_in; a.acts = acts_.ptr; a.dout16 = dout16_.ptr; a.dacts = dacts_.ptr;
        a.need_dx = n_grid_ ? 1 : 0;
        a.dx16 = (n_grid_ && keep_dx_) ? dx16_.ptr : nullptr;
        a.grid_grad = n_grid_ ? grad16_.ptr + n_mlp_ : nullptr;
        a.loss_partials = loss_partials_.ptr; a.loss_out = loss_dev_.ptr; a.n_loss_partials = tiles;
        if (train_groups_ > 0) {
            const uint32_t g = (uint32_t)train_groups_, grid2 = std::min<uint32_t>((tiles + g - 1) / g, (uint32_t)sm_count_ * 2);
            NRC_DISPATCH_INW(enc_.in_w, { nrc_backward2_kernel<IN_W><<<grid2, g * kGroupThreads, bwd2_smem_bytes<IN_W>(H), s>>>(a); });
        } else {
            NRC_DISPATCH_INW(enc_.in_w, { nrc_backward_kernel<IN_W><<<grid, threads, bwd_smem_bytes<IN_W>(H), s>>>(a); });
        }
        check_launch("nrc_backward_kernel");
        grid_grad_dirty_ = n_grid_ != 0;
    }
    {
        DwArgs a{};
        a.n_hidden = H; a.n = B; a.n_mlp = (uint32_t)n_mlp_;
        // enough chunks to fill the GPU for large batches (grid = chunks x (H + 1) CTAs), 32 for the reference's batch sizes
        const uint32_t want_chunks = tiles <= 1024 ? 32u : std::min<uint32_t>(kMaxDwChunks, tiles / 32);
        const uint32_t tiles_per_chunk = (tiles + want_chunks - 1) / want_chunks;
        a.kc = tiles_per_chunk * kTile;
        dw_chunks_ = (B + a.kc - 1) / a.kc;
        a.x16 = x16_.ptr; a.acts = acts_.ptr; a.dacts = dacts_.ptr; a.dout16 = dout16_.ptr; a.partials = dw_partials_.ptr;
        NRC_DISPATCH_INW(enc_.in_w, { nrc_dw_kernel<IN_W><<<dim3(dw_chunks_, H + 1), 128, kDwSmemBytes, s>>>(a); });
        check_launch("nrc_dw_kernel");
    }
}

void NrcCache::optimizer_step(cudaStream_t s) {
    NRCHPM_REQUIRE(grads_pending_, "nrc_optimizer_step without a preceding training step");
    current_step_++;
    OptArgs a{};
    a.n_params = n_params_; a.n_mlp = n_mlp_;
    a.master = master_.ptr; a.w16 = w16_.ptr; a.ema16 = ema16_.ptr; a.grad16 = grad16_.ptr; a.m1 = m1_.ptr; a.m2 = m2_.ptr; a.steps = steps_.ptr;
    a.grid_state = grid_state_.ptr;
    a.partials = dw_source_ ? dw_source_ : dw_partials_.ptr; a.n_chunks = dw_source_ ? 1u : dw_chunks_;
    a.lr = cfg_.learning_rate; a.beta1 = cfg_.beta1; a.beta2 = cfg_.beta2; a.eps = cfg_.epsilon; a.l2_reg = cfg_.l2_reg;
    a.loss_scale = cfg_.loss_scale * grad_scale_;          // peer_exchange leaves the SUM over the ranks in the buffers: divide here
    grad_scale_ = 1.0f;
    a.ema_decay = cfg_.ema_decay;
    a.log2_beta1 = std::log2(cfg_.beta1); a.log2_beta2 = std::log2(cfg_.beta2);
    a.ema_debias_old = 1 - (float)std::pow(cfg_.ema_decay, current_step_ - 1);          // ema.h:105-108
    a.ema_debias_new = 1.0f / (1 - (float)std::pow(cfg_.ema_decay, current_step_));
    a.tl = timeline_slot(); a.tl_ema = (n_params_ > n_mlp_) ? timeline_slot() : nullptr;
    NRCHPM_REQUIRE(n_mlp_ % 1024 == 0, "optimizer: the network parameter count must be a multiple of 1024");
    const bool has_grid = n_params_ > n_mlp_;
    if (has_grid && !ema_stream_) {
        NRCHPM_CUDA(cudaStreamCreateWithFlags(&ema_stream_, cudaStreamNonBlocking));
        NRCHPM_CUDA(cudaEventCreateWithFlags(&adam_done_, cudaEventDisableTiming));
        NRCHPM_CUDA(cudaEventCreateWithFlags(&ema_done_, cudaEventDisableTiming));
    }
    // ---- one launch: network weights (first n_mlp / 64 CTAs) + Adam on the touched hash-grid entries
    if (ema_in_flight_) NRCHPM_CUDA(cudaStreamWaitEvent(s, ema_done_, 0));        // the previous step's EMA pass still reads the weights
    a.mlp_blocks = (uint32_t)(n_mlp_ / 64);
    a.grid_begin = n_mlp_; a.grid_end = n_params_;
    const bool sharded = peer_sharded_step_ && has_grid;
    if (sharded) {                  // data-parallel replica: this rank updates its own slice of the encoding only (peer_exchange left the summed gradient there)
        uint64_t b, e;
        peer_slice(b, e);
        a.grid_begin = n_mlp_ + b * 8; a.grid_end = n_mlp_ + e * 8;
    }
    peer_sharded_step_ = false;
    const unsigned grid_blocks = has_grid ? (unsigned)(((a.grid_end - a.grid_begin) / 8 + 255) / 256) : 0u;
    if (sharded && peer_fused_) {
        // reduce-scatter + Adam on the slice + weight all-gather in ONE kernel over peer memory, then wait for every peer's final flag
        PeerArgs pa{};
        fill_peer_args(pa);
        cudaLaunchConfig_t cfg = {};
        // persistent grid over the slice's blocks of 256 words; shared memory (the cp.async ring) bounds the CTAs per SM
        const size_t smem = peer_world_ <= 2 ? peer_adam_smem_bytes<2>() : peer_world_ <= 4 ? peer_adam_smem_bytes<4>() : peer_adam_smem_bytes<8>();
        const unsigned per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (smem + 12 * 1024)));
        const unsigned peer_grid = std::min<unsigned>(grid_blocks, (peer_ctas_ ? peer_ctas_ : (unsigned)sm_count_ * per_sm));
        cfg.gridDim = a.mlp_blocks + peer_grid; cfg.blockDim = 256; cfg.stream = s; cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        if (l2_window_bytes_) {
            attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
            attr[0].val.accessPolicyWindow.base_ptr = hot_.ptr; attr[0].val.accessPolicyWindow.num_bytes = l2_window_bytes_;
            attr[0].val.accessPolicyWindow.hitRatio = l2_hit_ratio_; attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cfg.attrs = attr; cfg.numAttrs = 1;
        }
        if (peer_world_ <= 2) NRCHPM_CUDA(cudaLaunchKernelEx(&cfg, nrc_peer_adam_kernel<2>, a, pa));
        else if (peer_world_ <= 4) NRCHPM_CUDA(cudaLaunchKernelEx(&cfg, nrc_peer_adam_kernel<4>, a, pa));
        else NRCHPM_CUDA(cudaLaunchKernelEx(&cfg, nrc_peer_adam_kernel<8>, a, pa));
        check_launch("nrc_peer_adam_kernel");
        nrc_peer_wait_kernel<<<1, 32, 0, s>>>(peer_flag_words_.ptr, peer_world_, pa.token);
        check_launch("nrc_peer_wait_kernel");
    } else {
        launch_hot(nrc_adam_kernel, a.mlp_blocks + grid_blocks, 256, 0, s, a, true);
        check_launch("nrc_adam_kernel");
        if (sharded) peer_publish(s);   // ... and receives everybody else's
    }
    if (has_grid) {
        // ---- the dense EMA of the hash-grid weights, which only Inference() reads: side stream, underneath the next step.  (Round 2 also
        // measured it as extra CTAs of the next training launch, on the 20 SMs a 2^14 batch leaves idle: 20 SMs stream the 85 MB in
        // 85 us, longer than the training CTAs' 48 us -- 1.08 instead of 0.98 ms per bench step.)
        NRCHPM_CUDA(cudaEventRecord(adam_done_, s));
        launch_ema(a);
    }
    grid_grad_dirty_ = false;       // the optimizer re-zeroes every encoding gradient it consumed
    grads_pending_ = false;
}

// (Measured for the host path: launching the LAST step's pass behind the inference kernels, which read a snapshot and cannot share an SM
// with its 148 resident blocks, changes nothing -- 1.23-1.25 ms either way, profiles/r02_e2e_knobs.jsonl.)
void NrcCache::launch_ema(const OptArgs& a) {
    NRCHPM_CUDA(cudaStreamWaitEvent(ema_stream_, adam_done_, 0));
    if (ema_gate_) { NRCHPM_CUDA(cudaStreamWaitEvent(ema_stream_, ema_gate_, 0)); ema_gate_ = nullptr; }   // a snapshot copy still reads the EMA weights
    launch_hot(nrc_grid_ema_kernel, (unsigned)sm_count_, 256, 0, ema_stream_, a);
    check_launch("nrc_grid_ema_kernel");
    NRCHPM_CUDA(cudaEventRecord(ema_done_, ema_stream_));
    ema_in_flight_ = true;
}

// the EMA weights are complete once the last EMA pass has finished: every reader of ema16_ orders itself behind it
void NrcCache::wait_ema(cudaStream_t s) {
    if (ema_in_flight_) NRCHPM_CUDA(cudaStreamWaitEvent(s, ema_done_, 0));
}

float NrcCache::loss() {
    if (!loss_valid_) {
        cudaStream_t s = train_stream_last_;                              // the stream the last training step ran on
        NRCHPM_CUDA(cudaMemcpyAsync(&loss_host_, loss_dev_.ptr, sizeof(float), cudaMemcpyDeviceToHost, s));
        NRCHPM_CUDA(cudaStreamSynchronize(s));
        loss_valid_ = true;
    }
    return loss_host_;
}

// development aid (NRCHPM_TRAIN_PROF=1): clock64 stamps of the last fused training launch, [CTA][16]
uint32_t NrcCache::train_profile(long long* host_out, uint32_t max_ctas) {
    if (!train_prof_.ptr) return 0;
    NRCHPM_CUDA(cudaDeviceSynchronize());
    const uint32_t n = std::min<uint32_t>(max_ctas, (uint32_t)sm_count_);
    NRCHPM_CUDA(cudaMemcpy(host_out, train_prof_.ptr, (size_t)n * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    return n;
}
// development aid: every kernel of the training path stamps (earliest CTA start, latest CTA end) in %globaltimer ns into its own slot,
// in launch order: fused, adam, ema, fused, ... ; reading the log resets it
unsigned long long* NrcCache::timeline_slot() {
    if (!timeline_.ptr || tl_seq_ >= kTimelineSlots) return nullptr;
    return timeline_.ptr + 2 * (tl_seq_++);
}
void NrcCache::reset_timeline() {
    std::vector<unsigned long long> init(2 * kTimelineSlots);
    for (uint32_t k = 0; k < kTimelineSlots; k++) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
    NRCHPM_CUDA(cudaMemcpy(timeline_.ptr, init.data(), init.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    tl_seq_ = 0;
}
uint32_t NrcCache::read_timeline(unsigned long long* host_out, uint32_t max_slots) {
    if (!timeline_.ptr) return 0;
    NRCHPM_CUDA(cudaDeviceSynchronize());
    const uint32_t n = std::min(max_slots, tl_seq_);
    NRCHPM_CUDA(cudaMemcpy(host_out, timeline_.ptr, (size_t)n * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    reset_timeline();
    return n;
}

void NrcCache::last_step_tensor(int which, float* host_out) {
    NRCHPM_REQUIRE(last_batch_ > 0, "no training step has run");
    NRCHPM_CUDA(cudaDeviceSynchronize());
    const __half* src; size_t n;
    if (which == 0) { src = out16_.ptr; n = (size_t)last_batch_ * kOutPad; }
    else if (which == 1) { src = dout16_.ptr; n = (size_t)last_batch_ * kOutPad; }
    else if (which == 2) { NRCHPM_REQUIRE(n_grid_ && keep_dx_ && dx16_.ptr, "dL/dinput is only kept by the three-kernel training path (NRCHPM_TRAIN_FUSED=0) for encodings with parameters"); src = dx16_.ptr; n = (size_t)last_batch_ * enc_.in_w; }
    else throw Error(NRCHPM_ERR_INVALID, "nrc_last_step_tensor: which must be 0..2");
    std::vector<__half> h(n);
    NRCHPM_CUDA(cudaMemcpy(h.data(), src, n * sizeof(__half), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) host_out[i] = __half2float(h[i]);
}

// ------------------------------------------------------------------------------------------------ reference-level API
void NrcCache::init(uint32_t infer_count, float* in, float* out, float* train_in, float* train_target, cudaExternalSemaphore_t start,
                    cudaExternalSemaphore_t finished, cudaStream_t stream) {
    // src/NeuralRadianceCache.cu:52: inferCount % 16 == 0
    NRCHPM_REQUIRE(infer_count % 16 == 0, "NRC requires inferCount to be a multiple of 16");
    NRCHPM_REQUIRE(cfg_.train_batch_size % kTile == 0, "train batch size must be a multiple of 128");
    infer_count_ = infer_count; infer_in_ = in; infer_out_ = out; train_in_ = train_in; train_target_ = train_target;
    start_sem_ = start; finished_sem_ = finished; stream_ = stream;
    // src/NeuralRadianceCache.cu:66-80: floor(N/B) full batches + one remainder batch
    infer_batches_.clear();
    const uint32_t B = cfg_.infer_batch_size;
    for (uint32_t o = 0; o < infer_count; o += B) infer_batches_.push_back({o, std::min(B, infer_count - o)});
    ensure_train_scratch(cfg_.train_batch_size);
    initialised_ = true;
}

void NrcCache::run_inference(const uint32_t* filter_host) {
    for (size_t i = 0; i < infer_batches_.size(); i++) {
        if (filter_host && filter_host[i] == 0) continue;                        // src/NeuralRadianceCache.cu:136-144
        const auto& b = infer_batches_[i];
        inference(infer_in_ + 5 * (size_t)b.first, infer_out_ + 3 * (size_t)b.first, b.second, true, nullptr, nullptr, stream_);
    }
}

void NrcCache::run_train() {
    const uint32_t B = cfg_.train_batch_size;
    for (uint32_t i = 0; i < cfg_.train_batch_count; i++) {                      // src/NeuralRadianceCache.cu:147-156
        if (peer_world_ >= 2) {
            // data-parallel replica (nrc_peer_setup): every rank trains on its own tile's records and applies the mean gradient
            training_step(train_in_ + 5 * (size_t)i * B, train_target_ + 3 * (size_t)i * B, B, false, stream_);
            peer_exchange(stream_);
            optimizer_step(stream_);
        } else {
            training_step(train_in_ + 5 * (size_t)i * B, train_target_ + 3 * (size_t)i * B, B, true, stream_);
        }
    }
}

// Train() on caller-supplied record buffers and stream (pipelined frames: hpm_render with pipeline_train)
void NrcCache::run_train_on(const float* d_in, const float* d_target, cudaStream_t s) {
    const uint32_t B = cfg_.train_batch_size;
    for (uint32_t i = 0; i < cfg_.train_batch_count; i++) {
        if (peer_world_ >= 2) {
            training_step(d_in + 5 * (size_t)i * B, d_target + 3 * (size_t)i * B, B, false, s);
            peer_exchange(s);
            optimizer_step(s);
        } else {
            training_step(d_in + 5 * (size_t)i * B, d_target + 3 * (size_t)i * B, B, true, s);
        }
    }
}

void NrcCache::wait_start() {
    if (!start_sem_) return;
    cudaExternalSemaphoreWaitParams p{};
    NRCHPM_CUDA(cudaWaitExternalSemaphoresAsync(&start_sem_, &p, 1, stream_));   // src/NeuralRadianceCache.cu:158-167
}
void NrcCache::signal_finished() {
    if (!finished_sem_) return;
    cudaExternalSemaphoreSignalParams p{};
    NRCHPM_CUDA(cudaSignalExternalSemaphoresAsync(&finished_sem_, &p, 1, stream_));   // :169-178
}

void NrcCache::infer_and_train(const uint32_t* filter_host, bool train) {
    NRCHPM_REQUIRE(initialised_, "nrc_init has not been called");
    wait_start();
    bool all_batches = true;
    if (filter_host) for (size_t i = 0; i < infer_batches_.size(); i++) all_batches &= filter_host[i] != 0;
    if (train && peer_world_ >= 2 && overlap_schedule_ && all_batches && infer_count_ > 0) {
        infer_and_train_overlapped();
    } else {
        run_inference(filter_host);
        if (train) run_train();
    }
    signal_finished();
}

// Optional schedule (NRCHPM_OVERLAP=1) of InferAndTrain on a data-parallel replica (nrc_peer_setup).  It pays with the
// step-by-step exchange (NRCHPM_PEER_FUSED=0: ~50 us reduce-scatter + ~30 us all-gather windows per training step at two ranks); with
// the default fused peer kernel (nrc_peer_adam_kernel, ~40 us per step, SMs busy) the serial Inference() -> Train() order is faster
// (measured, profiles/r02_summary.md).  Every training step waits for the gradient exchange over
// NVLink, during which the SMs would idle; the frame's inference -- which in the reference's order runs first, on the weights of the
// previous frame -- is therefore cut into one chunk per training step and each chunk runs underneath that step's exchange + optimizer
// (stream s_inf) from a snapshot of the pre-training EMA weights, which keeps the reference's result.  The forward / backward kernel of
// the next step gets the whole GPU again: it waits for the chunk (it could not share an SM with the inference CTAs anyway: shared
// memory).  `overlap_head_` of the records can be evaluated up front at full occupancy when the windows are too short.
void NrcCache::infer_and_train_overlapped() {
    if (!ov_inf_stream_) {
        int prio_lo = 0, prio_hi = 0;
        NRCHPM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        NRCHPM_CUDA(cudaStreamCreateWithPriority(&ov_inf_stream_, cudaStreamNonBlocking, prio_lo));
        NRCHPM_CUDA(cudaStreamCreateWithPriority(&ov_tr_stream_, cudaStreamNonBlocking, prio_hi));
        for (auto& e : ov_ev_) NRCHPM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const uint32_t B = cfg_.train_batch_size, nb = cfg_.train_batch_count, n = infer_count_;
    NRCHPM_REQUIRE(nb <= kMaxOverlapBatches, "overlapped schedule: too many training batches per frame");
    cudaStream_t si = ov_inf_stream_, st = ov_tr_stream_;
    cudaEvent_t ev_begin = ov_ev_[0], ev_snap = ov_ev_[1], ev_inf_done = ov_ev_[2], ev_tr_done = ov_ev_[3];
    NRCHPM_CUDA(cudaEventRecord(ev_begin, stream_));
    NRCHPM_CUDA(cudaStreamWaitEvent(si, ev_begin, 0));
    NRCHPM_CUDA(cudaStreamWaitEvent(st, ev_begin, 0));
    // snapshot of the EMA weights Inference() has to see; only the first EMA pass of this frame's training overwrites them
    infer_snapshot_.ensure(n_params_);
    wait_ema(si);
    NRCHPM_CUDA(cudaMemcpyAsync(infer_snapshot_.ptr, ema16_.ptr, n_params_ * sizeof(__half), cudaMemcpyDeviceToDevice, si));
    NRCHPM_CUDA(cudaEventRecord(ev_snap, si));
    ema_gate_ = ev_snap;                                                // consumed by the next optimizer_step
    uint32_t off = 0;
    const uint32_t head = std::min(n, (uint32_t)((double)n * overlap_head_) / kTile * kTile);
    if (head) {
        inference_with(infer_snapshot_.ptr, infer_in_, infer_out_, head, nullptr, nullptr, si, 0);
        NRCHPM_CUDA(cudaEventRecord(ov_ev_[4], si));
        NRCHPM_CUDA(cudaStreamWaitEvent(st, ov_ev_[4], 0));
        off = head;
    }
    uint32_t chunk = (n - off + nb - 1) / nb;
    chunk = (chunk + kTile - 1) / kTile * kTile;
    peer_one_cta_per_sm_ = true;                                        // fits next to an inference CTA on every SM
    for (uint32_t b = 0; b < nb; b++) {
        cudaEvent_t ev_bwd = ov_ev_[5 + 2 * b], ev_chunk = ov_ev_[6 + 2 * b];
        training_step(train_in_ + 5 * (size_t)b * B, train_target_ + 3 * (size_t)b * B, B, false, st);
        NRCHPM_CUDA(cudaEventRecord(ev_bwd, st));
        peer_exchange(st);
        NRCHPM_CUDA(cudaStreamWaitEvent(si, ev_bwd, 0));
        const uint32_t m = b + 1 < nb ? std::min(chunk, n - off) : n - off;
        // the chunk takes `overlap_infer_sms_` SMs (one persistent CTA each, dispatched first: its stream is released by the backward pass,
        // the peer kernel follows a small reduction kernel); the peer kernel's CTAs cannot share an SM with an inference CTA (registers)
        // and fill the remaining SMs, where they keep the NVLink loads of ~4 CTAs per SM in flight
        if (m) inference_with(infer_snapshot_.ptr, infer_in_ + 5 * (size_t)off, infer_out_ + 3 * (size_t)off, m, nullptr, nullptr, si, overlap_infer_sms_);
        off += m;
        NRCHPM_CUDA(cudaEventRecord(ev_chunk, si));
        // the exchange kernels (one 48-register CTA per SM) and the optimizer on this rank's slice (small CTAs) share the SMs with the
        // chunk; the next forward / backward kernel (192 KB of shared memory per CTA) gets the whole GPU: it waits for the chunk
        optimizer_step(st);
        NRCHPM_CUDA(cudaStreamWaitEvent(st, ev_chunk, 0));
    }
    peer_one_cta_per_sm_ = false;
    NRCHPM_CUDA(cudaEventRecord(ev_inf_done, si));
    NRCHPM_CUDA(cudaEventRecord(ev_tr_done, st));
    NRCHPM_CUDA(cudaStreamWaitEvent(stream_, ev_inf_done, 0));
    NRCHPM_CUDA(cudaStreamWaitEvent(stream_, ev_tr_done, 0));
    train_stream_last_ = stream_;                                       // loss(): everything above is ordered before stream_ now
}

}  // namespace nrchpm

// ================================================================================================ C ABI
using namespace nrchpm;

extern "C" {

const char* nrchpm_last_error(void) { return t_last_error.c_str(); }
int nrchpm_version(void) { return 100; }
uint64_t nrchpm_launch_count(void) { return g_launch_count.load(); }

int nrc_create(const char* config_json, uint64_t seed, nrc_cache** out) {
    return guard([&] {
        NRCHPM_REQUIRE(config_json && out, "nrc_create: null argument");
        *out = nullptr;
        NrcConfig c = NrcConfig::from_json(config_json);
        *out = new nrc_cache(c, seed);
    });
}
int nrc_destroy(nrc_cache* c) { return guard([&] { delete c; }); }
int nrc_init(nrc_cache* c, uint32_t infer_count, float* in, float* out, float* tin, float* ttgt, void* s0, void* s1, void* stream) {
    return guard([&] {
        NRCHPM_REQUIRE(c, "null cache");
        c->impl.init(infer_count, in, out, tin, ttgt, (cudaExternalSemaphore_t)s0, (cudaExternalSemaphore_t)s1, (cudaStream_t)stream);
    });
}
int nrc_infer_and_train(nrc_cache* c, const uint32_t* f, int train) { return guard([&] { NRCHPM_REQUIRE(c, "null cache"); c->impl.infer_and_train(f, train != 0); }); }
int nrc_inference(nrc_cache* c, const uint32_t* f) {
    return guard([&] { NRCHPM_REQUIRE(c && c->impl.initialised(), "nrc_init has not been called"); c->impl.wait_start(); c->impl.run_inference(f); c->impl.signal_finished(); });
}
int nrc_train(nrc_cache* c) {
    return guard([&] { NRCHPM_REQUIRE(c && c->impl.initialised(), "nrc_init has not been called"); c->impl.wait_start(); c->impl.run_train(); c->impl.signal_finished(); });
}
int nrc_get_loss(nrc_cache* c, float* loss) { return guard([&] { NRCHPM_REQUIRE(c && loss, "null argument"); *loss = c->impl.loss(); }); }
size_t nrc_get_infer_batch_count(const nrc_cache* c) { return c ? c->impl.infer_batch_count() : 0; }
size_t nrc_get_train_batch_count(const nrc_cache* c) { return c ? c->impl.config().train_batch_count : 0; }
uint32_t nrc_get_infer_batch_size(const nrc_cache* c) { return c ? c->impl.config().infer_batch_size : 0; }
uint32_t nrc_get_train_batch_size(const nrc_cache* c) { return c ? c->impl.config().train_batch_size : 0; }
uint64_t nrc_n_params(const nrc_cache* c) { return c ? c->impl.n_params() : 0; }
uint64_t nrc_n_mlp_params(const nrc_cache* c) { return c ? c->impl.n_mlp() : 0; }
uint32_t nrc_input_width(const nrc_cache* c) { return c ? (uint32_t)c->impl.enc().in_w : 0; }
int nrc_get_params(nrc_cache* c, int which, float* out) { return guard([&] { NRCHPM_REQUIRE(c && out, "null argument"); c->impl.get_params(which, out); }); }
int nrc_set_params_fp32(nrc_cache* c, const float* m) { return guard([&] { NRCHPM_REQUIRE(c && m, "null argument"); c->impl.set_params_fp32(m); }); }
int nrc_set_ema(nrc_cache* c, const float* m) { return guard([&] { NRCHPM_REQUIRE(c && m, "null argument"); c->impl.set_ema(m); }); }
int nrc_gradient_buffers(nrc_cache* c, float** mlp, void** enc) {
    return guard([&] { NRCHPM_REQUIRE(c, "null cache"); c->impl.gradient_buffers(mlp, enc); });
}
int nrc_encode_batch(nrc_cache* c, const float* d_in, uint32_t n, int use_ema, void* d_out, void* stream) {
    return guard([&] { NRCHPM_REQUIRE(c && d_in && d_out, "null argument"); c->impl.encode(d_in, n, use_ema != 0, d_out, (cudaStream_t)stream); });
}
// ---- Vulkan interop (reference src/NrcHpmRenderer.cu:644-690, 700-821): import of exported buffers / semaphores by opaque fd
struct nrchpm_external_buffer { cudaExternalMemory_t mem = nullptr; void* ptr = nullptr; };
int nrchpm_import_external_buffer(int fd, size_t bytes, nrchpm_external_buffer** out, void** d_ptr_out) {
    return guard([&] {
        NRCHPM_REQUIRE(out && d_ptr_out, "null argument");
        *out = nullptr; *d_ptr_out = nullptr;
        NRCHPM_REQUIRE(fd >= 0 && bytes > 0, "external buffer: bad file descriptor or size");
        cudaExternalMemoryHandleDesc hd{};
        hd.type = cudaExternalMemoryHandleTypeOpaqueFd; hd.handle.fd = fd; hd.size = bytes;
        nrchpm_external_buffer b;
        NRCHPM_CUDA(cudaImportExternalMemory(&b.mem, &hd));
        cudaExternalMemoryBufferDesc bd{};
        bd.offset = 0; bd.size = bytes; bd.flags = 0;
        const cudaError_t e = cudaExternalMemoryGetMappedBuffer(&b.ptr, b.mem, &bd);
        if (e != cudaSuccess) { cudaDestroyExternalMemory(b.mem); NRCHPM_CUDA(e); }
        *out = new nrchpm_external_buffer(b); *d_ptr_out = b.ptr;
    });
}
int nrchpm_release_external_buffer(nrchpm_external_buffer* b) {
    return guard([&] {
        if (!b) return;
        if (b->ptr) cudaFree(b->ptr);                      // a mapped buffer is released with cudaFree, then the memory object
        if (b->mem) NRCHPM_CUDA(cudaDestroyExternalMemory(b->mem));
        delete b;
    });
}
int nrchpm_import_external_semaphore(int fd, void** sem_out) {
    return guard([&] {
        NRCHPM_REQUIRE(sem_out, "null argument");
        *sem_out = nullptr;
        NRCHPM_REQUIRE(fd >= 0, "external semaphore: bad file descriptor");
        cudaExternalSemaphoreHandleDesc hd{};
        hd.type = cudaExternalSemaphoreHandleTypeOpaqueFd; hd.handle.fd = fd;
        cudaExternalSemaphore_t sem = nullptr;
        NRCHPM_CUDA(cudaImportExternalSemaphore(&sem, &hd));
        *sem_out = (void*)sem;
    });
}
int nrchpm_release_external_semaphore(void* sem) {
    return guard([&] { if (sem) NRCHPM_CUDA(cudaDestroyExternalSemaphore((cudaExternalSemaphore_t)sem)); });
}

int nrc_inference_batch(nrc_cache* c, const float* d_in, float* d_out, uint32_t n, int use_ema, void* stream) {
    return guard([&] { NRCHPM_REQUIRE(c && d_in && d_out, "null argument"); c->impl.inference_set(use_ema, d_in, d_out, n, nullptr, nullptr, (cudaStream_t)stream); });
}
int nrc_inference_indexed(nrc_cache* c, const float* d_in, float* d_out, const uint32_t* idx, const uint32_t* cnt, uint32_t max_n, int use_ema, void* stream) {
    return guard([&] { NRCHPM_REQUIRE(c && d_in && d_out && idx, "null argument"); c->impl.inference_set(use_ema, d_in, d_out, max_n, idx, cnt, (cudaStream_t)stream); });
}
int nrc_peer_export(nrc_cache* c, uint8_t* handles_out) {
    return guard([&] { NRCHPM_REQUIRE(c && handles_out, "null argument"); c->impl.peer_export(handles_out); });
}
int nrc_peer_setup(nrc_cache* c, int rank, int world, const uint8_t* all_handles) {
    return guard([&] { NRCHPM_REQUIRE(c && all_handles, "null argument"); c->impl.peer_setup(rank, world, all_handles); });
}
int nrc_peer_exchange(nrc_cache* c, void* stream) {
    return guard([&] { NRCHPM_REQUIRE(c, "null cache"); c->impl.peer_exchange((cudaStream_t)stream); });
}
int nrc_set_inference_cta_limit(nrc_cache* c, uint32_t max_ctas) {
    return guard([&] { NRCHPM_REQUIRE(c, "null cache"); c->impl.set_inference_cta_limit(max_ctas); });
}
int nrc_snapshot_params(nrc_cache* c, int use_ema, void* stream) {
    return guard([&] { NRCHPM_REQUIRE(c, "null cache"); c->impl.snapshot_params(use_ema != 0, (cudaStream_t)stream); });
}
int nrc_training_step(nrc_cache* c, const float* d_in, const float* d_tgt, uint32_t batch, int run_opt, void* stream) {
    return guard([&] { NRCHPM_REQUIRE(c && d_in && d_tgt, "null argument"); c->impl.training_step(d_in, d_tgt, batch, run_opt != 0, (cudaStream_t)stream); });
}
int nrc_optimizer_step(nrc_cache* c, void* stream) { return guard([&] { NRCHPM_REQUIRE(c, "null cache"); c->impl.optimizer_step((cudaStream_t)stream); }); }
uint32_t nrc_debug_train_profile(nrc_cache* c, long long* out, uint32_t max_ctas) { uint32_t n = 0; guard([&] { NRCHPM_REQUIRE(c && out, "null argument"); n = c->impl.train_profile(out, max_ctas); }); return n; }
uint32_t nrc_debug_timeline(nrc_cache* c, unsigned long long* out, uint32_t max_slots) { uint32_t n = 0; guard([&] { NRCHPM_REQUIRE(c && out, "null argument"); n = c->impl.read_timeline(out, max_slots); }); return n; }
int nrc_last_step_tensor(nrc_cache* c, int which, float* out) { return guard([&] { NRCHPM_REQUIRE(c && out, "null argument"); c->impl.last_step_tensor(which, out); }); }

int nrc_inference_host(nrc_cache* c, const float* h_in, float* h_out, uint32_t n, int use_ema) {
    return guard([&] { NRCHPM_REQUIRE(c && h_in && h_out, "null argument"); c->impl.inference_host(h_in, h_out, n, use_ema != 0); });
}
int nrc_training_step_host(nrc_cache* c, const float* h_in, const float* h_tgt, uint32_t batch, float* loss_out) {
    return guard([&] { NRCHPM_REQUIRE(c && h_in && h_tgt, "null argument"); c->impl.training_step_host(h_in, h_tgt, batch, loss_out); });
}
int nrc_infer_and_train_host(nrc_cache* c, const float* h_in, float* h_out, uint32_t n, const float* h_tin, const float* h_tgt, uint32_t batch,
                             uint32_t n_batches, int use_ema, float* loss_out) {
    return guard([&] {
        NRCHPM_REQUIRE(c && (n == 0 || (h_in && h_out)) && (n_batches == 0 || (h_tin && h_tgt)), "null argument");
        c->impl.infer_and_train_host(h_in, h_out, n, h_tin, h_tgt, batch, n_batches, use_ema != 0, loss_out);
    });
}

}  // extern "C"

namespace nrchpm {
// Data-parallel training: collapse the per-chunk weight-gradient partials into one fp32 buffer that the caller can
// all-reduce (together with the fp16 encoding gradient) before nrc_optimizer_step consumes both.
void NrcCache::gradient_buffers(float** mlp, void** enc) {
    NRCHPM_REQUIRE(grads_pending_, "nrc_gradient_buffers: call nrc_training_step(run_optimizer=0) first");
    if (!dw_source_) {
        mlp_grad_f32_.ensure(n_mlp_);
        // on the stream of the training step that produced the partials (not the cache's own stream: the caller may train elsewhere)
        nrc_reduce_partials_kernel<<<(unsigned)((n_mlp_ + 63) / 64), 256, 0, train_stream_last_>>>(dw_partials_.ptr, dw_chunks_, (uint32_t)n_mlp_, mlp_grad_f32_.ptr);
        check_launch("nrc_reduce_partials_kernel");
        dw_source_ = mlp_grad_f32_.ptr;
    }
    if (mlp) *mlp = mlp_grad_f32_.ptr;
    if (enc) *enc = n_grid_ ? (void*)(grad16_.ptr + n_mlp_) : nullptr;
}
// ---- gradient exchange over peer memory (kernel: nrc_peer_reduce_kernel)
void NrcCache::peer_export(uint8_t* out) {
    NRCHPM_REQUIRE(n_grid_ % 8 == 0, "peer exchange needs the encoding gradient to be a multiple of 16 bytes");
    mlp_grad_f32_.ensure(n_mlp_); mlp_sum_.ensure(n_mlp_);
    if (!peer_flag_words_.ptr) { peer_flag_words_.allocate(3 * kMaxPeers); peer_flag_words_.zero(); peer_done_.allocate(1); peer_done_.zero(); }
    NRCHPM_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h[3];
    NRCHPM_CUDA(cudaIpcGetMemHandle(&h[0], grad16_.ptr));
    NRCHPM_CUDA(cudaIpcGetMemHandle(&h[1], mlp_grad_f32_.ptr));
    NRCHPM_CUDA(cudaIpcGetMemHandle(&h[2], peer_flag_words_.ptr));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    std::memcpy(out, h, sizeof(h));
}
void NrcCache::peer_setup(int rank, int world, const uint8_t* handles) {
    NRCHPM_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world, "peer exchange: 2..8 ranks");
    NRCHPM_REQUIRE(peer_flag_words_.ptr, "peer_setup before peer_export");
    for (int p = 0; p < world; p++) {
        if (p == rank) { peer_grad_[p] = grad16_.ptr; peer_mlp_[p] = mlp_grad_f32_.ptr; peer_flags_[p] = peer_flag_words_.ptr; continue; }
        cudaIpcMemHandle_t h[3];
        std::memcpy(h, handles + (size_t)p * kPeerHandleBytes, sizeof(h));
        NRCHPM_CUDA(cudaIpcOpenMemHandle(&peer_grad_[p], h[0], cudaIpcMemLazyEnablePeerAccess));
        NRCHPM_CUDA(cudaIpcOpenMemHandle(&peer_mlp_[p], h[1], cudaIpcMemLazyEnablePeerAccess));
        NRCHPM_CUDA(cudaIpcOpenMemHandle(&peer_flags_[p], h[2], cudaIpcMemLazyEnablePeerAccess));
    }
    peer_rank_ = rank; peer_world_ = world;
}
// own slice of the hash-grid entries, in int4 words (warp-aligned: 32 words = the 256 parameters one optimizer warp owns)
void NrcCache::peer_slice(uint64_t& begin, uint64_t& end) const {
    const uint64_t n_vec = n_grid_ / 8;
    uint64_t per = (n_vec + peer_world_ - 1) / peer_world_;
    per = (per + 31) / 32 * 32;
    begin = std::min(n_vec, per * (uint64_t)peer_rank_);
    end = std::min(n_vec, begin + per);
}
void NrcCache::fill_peer_args(PeerArgs& a) {
    a.rank = peer_rank_; a.world = peer_world_; a.token = peer_token_;
    const size_t n_pad = (n_params_ + 63) / 64 * 64;
    for (int p = 0; p < peer_world_; p++) {
        __half* base = reinterpret_cast<__half*>(peer_grad_[p]);                 // [grad16 | w16] of rank p
        a.grad[p] = reinterpret_cast<int4*>(base + n_mlp_);
        a.w16[p] = reinterpret_cast<int4*>(base + n_pad + n_mlp_);
        a.mlp[p] = reinterpret_cast<const float*>(peer_mlp_[p]);
        a.flags[p] = reinterpret_cast<uint32_t*>(peer_flags_[p]);
    }
    a.mlp_sum = mlp_sum_.ptr; a.n_vec = n_grid_ / 8; a.n_mlp = (uint32_t)n_mlp_; a.done_counter = peer_done_.ptr;
    peer_slice(a.slice_begin, a.slice_end);
}
// step 1 of the sharded optimizer step (nrc_peer_gather_kernel): afterwards this rank's gradient buffer holds the SUM over the ranks on
// its own slice and zeros elsewhere; nrc_optimizer_step then updates the slice and publishes the new weights to every peer
void NrcCache::peer_exchange(cudaStream_t s) {
    NRCHPM_REQUIRE(peer_world_ >= 2, "nrc_peer_exchange before nrc_peer_setup");
    NRCHPM_REQUIRE(grads_pending_ && !dw_source_, "nrc_peer_exchange: call nrc_training_step(run_optimizer=0) first");
    NRCHPM_REQUIRE(n_mlp_ % 8 == 0 && ((n_params_ + 63) / 64 * 64 + n_mlp_) % 8 == 0, "peer exchange: the encoding part of the fp16 vectors must be 16-byte aligned");
    // round A also tells the peers that this rank's previous EMA pass is done with the weights they are about to overwrite
    if (ema_in_flight_) NRCHPM_CUDA(cudaStreamWaitEvent(s, ema_done_, 0));
    nrc_reduce_partials_kernel<<<(unsigned)((n_mlp_ + 63) / 64), 256, 0, s>>>(dw_partials_.ptr, dw_chunks_, (uint32_t)n_mlp_, mlp_grad_f32_.ptr);
    check_launch("nrc_reduce_partials_kernel");
    ++peer_token_;
    if (!peer_fused_) {             // step-by-step form: reduce-scatter now, Adam on the slice and the weight all-gather in nrc_optimizer_step
        PeerArgs a{};
        fill_peer_args(a);
        a.tl = timeline_slot();
        const unsigned ctas = peer_ctas_ ? peer_ctas_ : (unsigned)sm_count_ * (peer_one_cta_per_sm_ ? 1 : 2);
        if (peer_world_ <= 2) launch_hot(nrc_peer_gather_kernel<2>, ctas, 256, 0, s, a);
        else if (peer_world_ <= 4) launch_hot(nrc_peer_gather_kernel<4>, ctas, 256, 0, s, a);
        else launch_hot(nrc_peer_gather_kernel<8>, ctas, 256, 0, s, a);
        check_launch("nrc_peer_gather_kernel");
    }
    dw_source_ = mlp_sum_.ptr;
    grad_scale_ = (float)peer_world_;
    peer_sharded_step_ = true;
}
// step 3: all-gather of the weight slices, then wait until every peer's slice has landed here
void NrcCache::peer_publish(cudaStream_t s) {
    PeerArgs a{};
    fill_peer_args(a);
    a.tl = timeline_slot();
    const unsigned ctas = peer_ctas_ ? peer_ctas_ : (unsigned)sm_count_ * (peer_one_cta_per_sm_ ? 1 : 2);
    launch_hot(nrc_peer_publish_kernel, ctas, 256, 0, s, a);
    check_launch("nrc_peer_publish_kernel");
    nrc_peer_wait_kernel<<<1, 32, 0, s>>>(peer_flag_words_.ptr, peer_world_, a.token);
    check_launch("nrc_peer_wait_kernel");
}

// Host-buffer inference: the records are cut into chunks of whole persistent-grid rounds and pipelined over three streams
// (H2D copy, compute, D2H copy), so PCIe transfers of chunk i+1 / i-1 overlap the kernel of chunk i.
void NrcCache::ensure_pipeline(uint32_t n_chunks) {
    if (!copy_in_stream_) {
        NRCHPM_CUDA(cudaStreamCreateWithFlags(&copy_in_stream_, cudaStreamNonBlocking));
        NRCHPM_CUDA(cudaStreamCreateWithFlags(&copy_out_stream_, cudaStreamNonBlocking));
        NRCHPM_CUDA(cudaStreamCreateWithFlags(&compute_stream_, cudaStreamNonBlocking));
        // training gets the higher priority: its (small) kernels slot in as soon as their records have landed -- i.e. while the first
        // inference chunk is still crossing PCIe -- instead of queueing behind the persistent inference launches
        int prio_lo = 0, prio_hi = 0;
        NRCHPM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        NRCHPM_CUDA(cudaStreamCreateWithPriority(&train_stream_, cudaStreamNonBlocking, prio_hi));
    }
    while (pipe_events_.size() < 2 * (size_t)n_chunks + 6) {
        cudaEvent_t e; NRCHPM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        pipe_events_.push_back(e);
    }
}
// queues the chunked H2D -> kernel -> D2H pipeline; returns without waiting.  `off` holds the chunk boundaries (record indices, off[0] = 0,
// off.back() = n); copies_queued: the H2D copies and their events are already on copy_in_stream_ (queue_inference_copies_in)
// development aid (NRCHPM_E2E_TRACE=1): timing events at the hand-over points of the host pipeline, printed to stderr as one JSON line per call
void NrcCache::trace_mark(const char* name, cudaStream_t s) {
    if (!trace_on_) return;
    cudaEvent_t e; NRCHPM_CUDA(cudaEventCreate(&e)); NRCHPM_CUDA(cudaEventRecord(e, s));
    trace_.emplace_back(name, e);
}
static std::vector<uint32_t> uniform_chunks(uint32_t n, uint32_t chunk) {
    std::vector<uint32_t> off{0u};
    while (off.back() < n) off.push_back(std::min(n, off.back() + chunk));
    return off;
}
void NrcCache::queue_inference_copies_in(const float* h_in, const std::vector<uint32_t>& off) {
    for (size_t c = 0; c + 1 < off.size(); c++) {
        const uint32_t o = off[c], m = off[c + 1] - o;
        NRCHPM_CUDA(cudaMemcpyAsync(host_in_.ptr + (size_t)o * 5, h_in + (size_t)o * 5, (size_t)m * 5 * sizeof(float), cudaMemcpyHostToDevice, copy_in_stream_));
        NRCHPM_CUDA(cudaEventRecord(pipe_events_[2 * c], copy_in_stream_));
        trace_mark("h2d", copy_in_stream_);
    }
}
void NrcCache::queue_inference_pipeline(const __half* params, const float* h_in, float* h_out, const std::vector<uint32_t>& off, bool copies_queued) {
    if (!copies_queued) queue_inference_copies_in(h_in, off);
    for (size_t c = 0; c + 1 < off.size(); c++) {
        const uint32_t o = off[c], m = off[c + 1] - o;
        NRCHPM_CUDA(cudaStreamWaitEvent(compute_stream_, pipe_events_[2 * c], 0));
        trace_mark("inf_begin", compute_stream_);
        inference_with(params, host_in_.ptr + (size_t)o * 5, host_out_.ptr + (size_t)o * 3, m, nullptr, nullptr, compute_stream_, 0);
        NRCHPM_CUDA(cudaEventRecord(pipe_events_[2 * c + 1], compute_stream_));
        trace_mark("inf_end", compute_stream_);
        NRCHPM_CUDA(cudaStreamWaitEvent(copy_out_stream_, pipe_events_[2 * c + 1], 0));
        NRCHPM_CUDA(cudaMemcpyAsync(h_out + (size_t)o * 3, host_out_.ptr + (size_t)o * 3, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, copy_out_stream_));
        trace_mark("d2h", copy_out_stream_);
    }
}
void NrcCache::inference_host(const float* h_in, float* h_out, uint32_t n, bool use_ema) {
    if (n == 0) return;
    host_in_.ensure((size_t)n * 5); host_out_.ensure((size_t)n * 3);
    const uint32_t chunk = (uint32_t)sm_count_ * 2 * 2 * 4 * kTile;       // 4 tiles per resident warpgroup
    const std::vector<uint32_t> off = uniform_chunks(n, chunk);
    const uint32_t n_chunks = (uint32_t)off.size() - 1;
    ensure_pipeline(n_chunks);
    // everything already queued on the cache's stream (e.g. a training step that changed the weights) comes first
    cudaEvent_t ev_prev = pipe_events_[2 * (size_t)n_chunks];
    NRCHPM_CUDA(cudaEventRecord(ev_prev, stream_));
    NRCHPM_CUDA(cudaStreamWaitEvent(compute_stream_, ev_prev, 0));
    NRCHPM_CUDA(cudaStreamWaitEvent(copy_in_stream_, ev_prev, 0));
    if (use_ema) wait_ema(compute_stream_);
    queue_inference_pipeline(use_ema ? ema16_.ptr : w16_.ptr, h_in, h_out, off, false);
    NRCHPM_CUDA(cudaStreamSynchronize(copy_out_stream_));
}
void NrcCache::training_step_host(const float* h_in, const float* h_tgt, uint32_t B, float* loss_out) {
    host_tin_.ensure((size_t)B * 5); host_tgt_.ensure((size_t)B * 3);
    cudaStream_t s = stream_;
    NRCHPM_CUDA(cudaMemcpyAsync(host_tin_.ptr, h_in, (size_t)B * 5 * sizeof(float), cudaMemcpyHostToDevice, s));
    NRCHPM_CUDA(cudaMemcpyAsync(host_tgt_.ptr, h_tgt, (size_t)B * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    training_step(host_tin_.ptr, host_tgt_.ptr, B, true, s);
    if (loss_out) *loss_out = loss();
}
// en::NeuralRadianceCache::InferAndTrain (src/NeuralRadianceCache.cu:97-156) on HOST buffers in ONE call.  The reference runs
// Inference() with the parameters of the previous frame and then Train(); the result of that order is kept, but not the serial
// schedule: the inference pipeline (PCIe-bound: 20 B in + 12 B out per record) reads a device-to-device SNAPSHOT of the
// pre-training parameters (28.5 MB, ~10 us), so the training steps -- whose records travel first, they are small -- run on their
// own stream underneath the record transfers instead of behind them.  The host waits once, at the end.
void NrcCache::infer_and_train_host(const float* h_in, float* h_out, uint32_t n, const float* h_tin, const float* h_tgt, uint32_t B,
                                    uint32_t n_batches, bool use_ema, float* loss_out) {
    const bool train = n_batches > 0 && B > 0;
    if (train) NRCHPM_REQUIRE(B % kTile == 0, "training batch must be a positive multiple of 128");
    if (n) { host_in_.ensure((size_t)n * 5); host_out_.ensure((size_t)n * 3); }
    const size_t T = (size_t)B * n_batches;
    if (train) { host_tin_.ensure(T * 5); host_tgt_.ensure(T * 3); ensure_train_scratch(B); }
    // 4 chunks for a 1080p frame (8 tiles per resident warpgroup).  Measured alternatives (profiles/r02_e2e_knobs.jsonl): 8 / 16 uniform
    // chunks 1.29 / 1.38 ms, a large first chunk and shrinking later ones (40/25/17/11/7 %) 1.25, two or three chunks 1.27-1.45, against 1.24;
    // the snapshot the kernels read kept in the persisting part of L2 (access-policy window on these launches): 1.25 against 1.245;
    // ONE persistent inference launch per frame whose producers wait for per-chunk flags raised behind the H2D copies and whose
    // consumers count finished tiles for the D2H stream (cuStreamWriteValue32 / cuStreamWaitValue32): 1.235 against 1.25 -- correct in
    // the parity tests, but 1 % does not pay for an ordering that rests on fences instead of stream events.
    uint32_t chunk = (uint32_t)sm_count_ * 2 * 2 * 8 * kTile;
    if (const char* v = std::getenv("NRCHPM_E2E_CHUNK_TILES")) chunk = (uint32_t)sm_count_ * 2 * 2 * (uint32_t)std::max(1, std::atoi(v)) * kTile;   // experiment knob
    const std::vector<uint32_t> off = uniform_chunks(n, chunk);
    const uint32_t n_chunks = (uint32_t)off.size() - 1;
    ensure_pipeline(n_chunks);
    const size_t e0 = 2 * (size_t)n_chunks;
    cudaEvent_t ev_prev = pipe_events_[e0], ev_train = pipe_events_[e0 + 1], ev_done = pipe_events_[e0 + 2], ev_snap = pipe_events_[e0 + 3], ev_trained = pipe_events_[e0 + 4], ev_loss = pipe_events_[e0 + 5];
    const bool overlap = train && n > 0;
    cudaStream_t ts = overlap ? train_stream_ : compute_stream_;
    static const bool trace_env = std::getenv("NRCHPM_E2E_TRACE") != nullptr;
    trace_on_ = trace_env;
    const auto host_t0 = std::chrono::steady_clock::now();
    NRCHPM_CUDA(cudaEventRecord(ev_prev, stream_));
    NRCHPM_CUDA(cudaStreamWaitEvent(compute_stream_, ev_prev, 0));
    NRCHPM_CUDA(cudaStreamWaitEvent(copy_in_stream_, ev_prev, 0));
    trace_mark("start", copy_in_stream_);
    if (overlap) {
        infer_snapshot_.ensure(n_params_);
        if (use_ema) wait_ema(compute_stream_);
        NRCHPM_CUDA(cudaMemcpyAsync(infer_snapshot_.ptr, use_ema ? ema16_.ptr : w16_.ptr, n_params_ * sizeof(__half), cudaMemcpyDeviceToDevice, compute_stream_));
        NRCHPM_CUDA(cudaEventRecord(ev_snap, compute_stream_));
        trace_mark("snapshot", compute_stream_);
        NRCHPM_CUDA(cudaStreamWaitEvent(train_stream_, ev_snap, 0));        // training overwrites what the snapshot copy reads
    }
    if (train) {
        NRCHPM_CUDA(cudaMemcpyAsync(host_tin_.ptr, h_tin, T * 5 * sizeof(float), cudaMemcpyHostToDevice, copy_in_stream_));
        NRCHPM_CUDA(cudaMemcpyAsync(host_tgt_.ptr, h_tgt, T * 3 * sizeof(float), cudaMemcpyHostToDevice, copy_in_stream_));
        NRCHPM_CUDA(cudaEventRecord(ev_train, copy_in_stream_));
        trace_mark("train_h2d", copy_in_stream_);
    }
    // PCIe is the scarce resource of this call (20 B in + 12 B out per record): every H2D copy is queued before the host spends its
    // time on the training launches
    if (n) queue_inference_copies_in(h_in, off);
    // Training goes FIRST on the device: its records are small and have landed long before the first inference chunk has crossed
    // PCIe, and the persistent inference launches cannot share an SM with the training kernels (TMEM and registers), so interleaving
    // the two only makes every training kernel wait for a whole chunk to drain (measured: 1.81 ms; training first: see profiles/;
    // inference chunks capped to 88..124 SMs NEXT TO the training steps: 1.31-1.52 ms against 1.26, profiles/r02_e2e_knobs.jsonl).
    // The inference kernels wait for the last optimizer step; their H2D copies do not.  Inference still evaluates the snapshot.
    if (train) {
        NRCHPM_CUDA(cudaStreamWaitEvent(ts, ev_train, 0));
        trace_mark("train_begin", ts);
        for (uint32_t b = 0; b < n_batches; b++) {
            if (b) trace_mark("train_step", ts);
            if (peer_world_ >= 2) {      // data-parallel replica: mean gradient over the ranks
                training_step(host_tin_.ptr + (size_t)b * B * 5, host_tgt_.ptr + (size_t)b * B * 3, B, false, ts);
                peer_exchange(ts);
                optimizer_step(ts);
            } else {
                training_step(host_tin_.ptr + (size_t)b * B * 5, host_tgt_.ptr + (size_t)b * B * 3, B, true, ts);
            }
        }
        trace_mark("train_end", ts);
        if (overlap) {
            // recorded BEFORE the loss read-back: the inference kernels wait for the optimizer, not for a copy-engine round trip
            NRCHPM_CUDA(cudaEventRecord(ev_trained, train_stream_));
            NRCHPM_CUDA(cudaStreamWaitEvent(compute_stream_, ev_trained, 0));
        }
        if (!loss_pinned_) NRCHPM_CUDA(cudaMallocHost((void**)&loss_pinned_, sizeof(float)));   // pageable memory would make this copy block the host
        NRCHPM_CUDA(cudaMemcpyAsync(loss_pinned_, loss_dev_.ptr, sizeof(float), cudaMemcpyDeviceToHost, ts));
        if (overlap) {
            NRCHPM_CUDA(cudaEventRecord(ev_loss, train_stream_));
            NRCHPM_CUDA(cudaStreamWaitEvent(copy_out_stream_, ev_loss, 0));                    // the host's final wait on copy_out_stream_ covers the loss
        }
    }
    if (n) {
        if (!overlap && use_ema) wait_ema(compute_stream_);
        queue_inference_pipeline(overlap ? infer_snapshot_.ptr : (use_ema ? ema16_.ptr : w16_.ptr), h_in, h_out, off, true);
    }
    // later work on the cache's own stream is ordered behind this call
    NRCHPM_CUDA(cudaEventRecord(ev_done, compute_stream_));
    NRCHPM_CUDA(cudaStreamWaitEvent(stream_, ev_done, 0));
    const auto host_t1 = std::chrono::steady_clock::now();
    NRCHPM_CUDA(cudaStreamSynchronize(compute_stream_));
    NRCHPM_CUDA(cudaStreamSynchronize(copy_out_stream_));
    if (train) { loss_host_ = *loss_pinned_; loss_valid_ = true; if (loss_out) *loss_out = loss_host_; }
    if (trace_on_) {
        const auto host_t2 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "{\"e2e_trace_us\": {\"host_enqueue\": %.1f, \"host_total\": %.1f, \"events\": [", std::chrono::duration<double, std::micro>(host_t1 - host_t0).count(),
                     std::chrono::duration<double, std::micro>(host_t2 - host_t0).count());
        for (size_t i = 0; i < trace_.size(); i++) {
            float ms = 0; cudaEventSynchronize(trace_[i].second); cudaEventElapsedTime(&ms, trace_[0].second, trace_[i].second);
            std::fprintf(stderr, "%s[\"%s\", %.1f]", i ? ", " : "", trace_[i].first.c_str(), ms * 1e3f);
        }
        std::fprintf(stderr, "]}}\n");
        for (auto& t : trace_) cudaEventDestroy(t.second);
        trace_.clear(); trace_on_ = false;
    }
}
}  // namespace nrchpm
