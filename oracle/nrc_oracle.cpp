// TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle "A") of the Neural Radiance Cache arithmetic
// that the reference delegates to tiny-cuda-nn (tcnn @6f018a9, vendored submodule).  Nothing under
// oracle/ is linked, imported or executed by the product path; only tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs use it.
//
// Restated pieces (paths relative to /root/reference/tiny-cuda-nn unless noted):
//   src/NeuralRadianceCache.cu:11-40 (reference repo)  model JSON: EMA(Adam), Composite encoding, FullyFusedMLP
//   include/tiny-cuda-nn/encodings/grid.h:49-212, 215-320, 652-724     hash grid fwd / bwd / offsets table
//   include/tiny-cuda-nn/common_device.h:632-718, 842-868, 905-920      hashes, grid_index, pos_fract, quartic_cdf
//   include/tiny-cuda-nn/encodings/oneblob.h:46-127, 183-229            OneBlob AoS + SoA kernels (incl. SoA pad bug, Q6)
//   include/tiny-cuda-nn/encodings/{triangle_wave.h:46-82,frequency.h:46-80,identity.h:46-66}
//   include/tiny-cuda-nn/encodings/composite.h:137-218                   nested widths / padding
//   src/fully_fused_mlp.cu:47-129, 150-259, 636-836, 866-891            MLP fwd / bwd / weight grads / init
//   include/tiny-cuda-nn/losses/relative_l2_luminance.h:40-88
//   include/tiny-cuda-nn/optimizers/adam.h:48-121, optimizers/ema.h:63-76,102-138
//   include/tiny-cuda-nn/trainer.h:50-87,163-211, random.h:40-66, gpu_matrix.h:284-299, dependencies/pcg32/pcg32.h
//
// Parity status: the reference ships no golden vectors for this arithmetic (tcnn has no tests).  The
// oracle is pinned against outputs of the REAL tcnn built for sm_100a (oracle/tcnn_ref) on a B200:
// initial parameters bit-for-bit, network_input / outputs / gradients / loss curves to tolerance;
// the committed fixtures are under tests/golden/tcnn_*.npz with the generating script.
//
// fp16 values are carried as floats that are exactly representable in binary16.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>
#include <algorithm>

extern "C" {
struct NrcoConfig {
    int32_t pos_enc;          // 0 HashGrid 1 Identity 2 TriangleWave 3 Frequency   (reference src/AppConfig.cpp:16-48)
    int32_t dir_enc;          // 0 OneBlob  1 Identity 2 TriangleWave               (reference src/AppConfig.cpp:50-73)
    int32_t n_neurons;        // 64
    int32_t n_hidden_layers;  // nnDepth
    int32_t oneblob_soa_bug;  // 1: reproduce tcnn's SoA padding bug (rows 34..41 = 1, 42..47 = 0); only with pos_enc==0 && dir_enc==0
    int32_t accum_fp16;       // 1: emulate wmma fp16 accumulators per k=16 block (tcnn, Q8); 0: fp32 accumulate
    int32_t n_levels, log2_hashmap_size, base_resolution;
    float per_level_scale;
    int32_t n_frequencies_pos, n_frequencies_dir, n_bins;
    float learning_rate, ema_decay, beta1, beta2, epsilon, l2_reg, loss_scale;
};
}

namespace {

inline float h(float x) { return (float)(_Float16)x; }
inline float hd(double x) { return (float)(_Float16)x; }
inline float hfma(float a, float b, float c) { return hd((double)a * (double)b + (double)c); }   // single-rounded fp16 fma

struct Pcg32 {
    uint64_t state, inc;
    explicit Pcg32(uint64_t initstate, uint64_t initseq = 1u) {
        state = 0u; inc = (initseq << 1u) | 1u; next_uint(); state += initstate; next_uint();
    }
    uint32_t next_uint() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
    }
    float next_float() {
        uint32_t u = (next_uint() >> 9) | 0x3f800000u;
        float f; std::memcpy(&f, &u, 4);
        return f - 1.0f;
    }
    void advance(uint64_t delta) {
        uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
        while (delta > 0) {
            if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
            cur_plus = (cur_mult + 1) * cur_plus; cur_mult *= cur_mult; delta /= 2;
        }
        state = acc_mult * state + acc_plus;
    }
};

constexpr float TCNN_PI = 3.14159265358979323846f;

struct Model {
    NrcoConfig c;
    // derived
    int pos_w, dir_w, in_w, dir_off, out_pad;          // widths; dir_off = first row of the direction slice
    std::vector<uint32_t> offsets;                      // grid offset table (entries)
    std::vector<int> mat_rows, mat_cols; std::vector<size_t> mat_off;
    size_t n_mlp, n_grid, n_params;
    // state (tcnn Trainer buffers + optimizer state)
    std::vector<float> master, w16, g16, ema16, m1, m2;
    std::vector<uint32_t> steps;
    uint32_t current_step = 0;
    // scratch kept from the last training step for inspection
    std::vector<float> last_dL_dinput, last_output, last_dL_doutput;
};

float grid_scale(uint32_t level, float log2_pls, uint32_t base) { return exp2f((float)level * log2_pls) * (float)base - 1.0f; }
uint32_t grid_resolution(float scale) { return (uint32_t)ceilf(scale) + 1; }

void derive(Model& m) {
    const NrcoConfig& c = m.c;
    const int pw[4] = {c.n_levels * 2, 3, 3 * c.n_frequencies_pos, 3 * c.n_frequencies_pos * 2};
    const int dw[3] = {2 * c.n_bins, 2, 2 * c.n_frequencies_dir};
    m.pos_w = pw[c.pos_enc]; m.dir_w = dw[c.dir_enc];
    m.dir_off = m.pos_w;                                   // every dir encoding has required_output_alignment()==1
    m.in_w = ((m.pos_w + m.dir_w + 15) / 16) * 16;         // set_alignment(16), network_with_input_encoding.h:47
    m.out_pad = 16;                                        // next_multiple(3, 16), fully_fused_mlp.cu:656
    m.offsets.clear();
    m.n_grid = 0;
    if (c.pos_enc == 0) {
        uint32_t offset = 0;
        const float log2_pls = std::log2(c.per_level_scale);
        for (int i = 0; i < c.n_levels; i++) {
            const uint32_t res = grid_resolution(grid_scale(i, log2_pls, c.base_resolution));
            const uint32_t max_params = 0xFFFFFFFFu / 2;
            uint32_t p = std::pow((float)res, 3) > (float)max_params ? max_params : res * res * res;
            p = ((p + 7) / 8) * 8;
            p = std::min(p, 1u << c.log2_hashmap_size);
            m.offsets.push_back(offset);
            offset += p;
        }
        m.offsets.push_back(offset);
        m.n_grid = (size_t)offset * 2;
    }
    const int W = c.n_neurons;
    m.mat_rows.clear(); m.mat_cols.clear(); m.mat_off.clear();
    size_t off = 0;
    auto add = [&](int r, int cc) { m.mat_rows.push_back(r); m.mat_cols.push_back(cc); m.mat_off.push_back(off); off += (size_t)r * cc; };
    add(W, m.in_w);
    for (int i = 0; i < c.n_hidden_layers - 1; i++) add(W, W);
    add(m.out_pad, W);
    m.n_mlp = off;
    m.n_params = m.n_mlp + m.n_grid;     // [network | encoding], network_with_input_encoding.h:115-130
}

void init_params(Model& m, uint32_t seed) {
    std::seed_seq seq{seed};
    std::vector<uint32_t> seeds(2);
    seq.generate(seeds.begin(), seeds.end());
    Pcg32 rng{seeds.front()};
    // MLP: xavier uniform, matrix by matrix, row-major (gpu_matrix.h:284-299)
    for (size_t k = 0; k < m.mat_rows.size(); k++) {
        const float scale = std::sqrt(6.0f / (float)(m.mat_cols[k] + m.mat_rows[k]));
        float* w = m.master.data() + m.mat_off[k];
        const size_t n = (size_t)m.mat_rows[k] * m.mat_cols[k];
        for (size_t i = 0; i < n; i++) w[i] = rng.next_float() * 2.0f * scale - scale;
    }
    // grid: device-side generate_random_uniform (random.h:40-66), element i + n_threads*j <- float #(4i + j)
    if (m.n_grid) {
        const size_t n = m.n_grid;
        const size_t n_threads_req = (n + 3) / 4;
        const size_t n_threads = 128 * ((n_threads_req + 127) / 128);
        float* g = m.master.data() + m.n_mlp;
        const float lower = -1e-4f, upper = 1e-4f;
        Pcg32 base = rng;
        for (size_t i = 0; i < n_threads; i++) {
            Pcg32 r = base;          // sequential walk == advance(4*i)
            for (int j = 0; j < 4; j++) {
                const size_t idx = i + n_threads * j;
                const float v = r.next_float();
                if (idx < n) g[idx] = fmaf(v, upper - lower, lower);
            }
            base.next_uint(); base.next_uint(); base.next_uint(); base.next_uint();
        }
    }
    for (size_t i = 0; i < m.n_params; i++) m.w16[i] = h(m.master[i]);
}

// ---------------- encodings ----------------
float quartic_cdf(float x, float inv_radius) {
    const float u = x * inv_radius, u2 = u * u, u4 = u2 * u2;
    return fmaxf(0.0f, fminf(1.0f, ((float)15 / 16) * u * (1 - ((float)2 / 3) * u2 + ((float)1 / 5) * u4) + 0.5f));
}

uint32_t grid_index(uint32_t hashmap_size, uint32_t res, const uint32_t p[3]) {
    uint32_t stride = 1, index = 0;
    for (uint32_t dim = 0; dim < 3 && stride <= hashmap_size; ++dim) { index += p[dim] * stride; stride *= res; }
    if (hashmap_size < stride) index = (p[0] * 1u) ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u);
    return index % hashmap_size;
}

struct GridCell { uint32_t idx[8]; float w[8]; };
void grid_cell(const Model& m, int level, const float* pos3, GridCell* out) {
    const float log2_pls = std::log2(m.c.per_level_scale);
    const float scale = grid_scale(level, log2_pls, m.c.base_resolution);
    const uint32_t res = grid_resolution(scale);
    const uint32_t hashmap_size = m.offsets[level + 1] - m.offsets[level];
    float pos[3]; uint32_t pg[3];
    for (int d = 0; d < 3; d++) {
        float p = fmaf(scale, pos3[d], 0.5f);
        float tmp = floorf(p);
        // (uint32_t)(int)tmp: CUDA float->int saturates; keep defined behaviour on the host
        double td = tmp; if (!(td == td)) td = 0; td = std::min(std::max(td, -2147483648.0), 2147483647.0);
        pg[d] = (uint32_t)(int32_t)td;
        pos[d] = p - tmp;
    }
    for (uint32_t idx = 0; idx < 8; idx++) {
        float w = 1; uint32_t pl[3];
        for (uint32_t d = 0; d < 3; d++) {
            if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; } else { w *= pos[d]; pl[d] = pg[d] + 1; }
        }
        out->idx[idx] = (m.offsets[level] + grid_index(hashmap_size, res, pl)) * 2;
        out->w[idx] = w;
    }
}

void oneblob_soa(float x, int n_bins, float* out) {                   // oneblob.h:99-127
    const float nb = (float)n_bins;
    float left = quartic_cdf(-x, nb) + quartic_cdf(-x - 1.0f, nb) + quartic_cdf(-x + 1.0f, nb);
    for (int k = 0; k < n_bins; k++) {
        const float rb = (float)(k + 1) / nb;                          // scalbnf(k+1, -log2 n_bins)
        const float right = quartic_cdf(rb - x, nb) + quartic_cdf(rb - x - 1.0f, nb) + quartic_cdf(rb - x + 1.0f, nb);
        out[k] = h(right - left);
        left = right;
    }
}
void oneblob_aos(float x, int n_bins, float* out) {                   // oneblob.h:46-70, 85-97
    const float nb = (float)n_bins;
    std::vector<float> left(n_bins);
    for (int k = 0; k < n_bins; k++) {
        const float lb = (float)k / nb;
        left[k] = quartic_cdf(lb - x, nb) + quartic_cdf(lb - x - 1.0f, nb) + quartic_cdf(lb - x + 1.0f, nb);
    }
    for (int k = 0; k < n_bins; k++) {
        float right = left[(k + 1) % n_bins];
        if (k == n_bins - 1) right += 1;
        out[k] = h(right - left[k]);
    }
}
void triangle_wave(float x_in, int n_freq, float* out) {              // triangle_wave.h:46-82
    for (int f = 0; f < n_freq; f++) {
        const float x = scalbnf(x_in, f - 1);
        const float val = x + (float)f * 0.25f;
        out[f] = h(fabsf(val - floorf(val) - 0.5f) * 4 - 1);
    }
}

// params16: fp16-as-float parameter vector to use (working or EMA).  x: [in_w]
void encode(const Model& m, const float* params16, const float* rec, float* x) {
    const NrcoConfig& c = m.c;
    for (int i = 0; i < m.in_w; i++) x[i] = 1.0f;                                   // padding value
    // --- position ---
    if (c.pos_enc == 0) {
        const float* grid = params16 + m.n_mlp;
        for (int l = 0; l < c.n_levels; l++) {
            GridCell cell; grid_cell(m, l, rec, &cell);
            float r0 = 0, r1 = 0;
            for (int k = 0; k < 8; k++) {
                const float w = h(cell.w[k]);
                r0 = hfma(w, grid[cell.idx[k] + 0], r0);
                r1 = hfma(w, grid[cell.idx[k] + 1], r1);
            }
            x[2 * l] = r0; x[2 * l + 1] = r1;
        }
    } else if (c.pos_enc == 1) {
        for (int d = 0; d < 3; d++) x[d] = h(rec[d] * 1.0f + 0.0f);
    } else if (c.pos_enc == 2) {
        for (int d = 0; d < 3; d++) triangle_wave(rec[d], c.n_frequencies_pos, x + d * c.n_frequencies_pos);
    } else {
        const int nf = c.n_frequencies_pos;                                          // frequency.h:46-80 (tcnn uses __sinf)
        for (int j = 0; j < 3 * nf * 2; j++) {
            const int d = j / (nf * 2), l2f = (j / 2) % nf;
            const float phase = (float)(j % 2) * (TCNN_PI / 2);
            const float xx = scalbnf(rec[d], l2f);
            x[j] = h(sinf(xx * TCNN_PI + phase));
        }
    }
    // --- direction ---
    float* xd = x + m.dir_off;
    if (c.dir_enc == 0) {
        const bool soa = (c.pos_enc == 0);                     // composite.h:400-403: layout of the first nested encoding
        for (int d = 0; d < 2; d++) (soa ? oneblob_soa : oneblob_aos)(rec[3 + d], c.n_bins, xd + d * c.n_bins);
        if (soa && c.oneblob_soa_bug) {
            // oneblob.h:224-227: pad write starts at row n_dims_to_encode (=2) of the slice instead of
            // row n_output_dims (=8): rows 2..2+n_pad-1 := 1.0, the real pad rows stay uninitialised (-> 0 here).
            const int n_pad = m.in_w - m.dir_off - m.dir_w;
            for (int r = m.dir_w; r < m.dir_w + n_pad; r++) xd[r] = 0.0f;
            for (int r = 2; r < 2 + n_pad; r++) xd[r] = 1.0f;
        }
    } else if (c.dir_enc == 1) {
        for (int d = 0; d < 2; d++) xd[d] = h(rec[3 + d]);
    } else {
        for (int d = 0; d < 2; d++) triangle_wave(rec[3 + d], c.n_frequencies_dir, xd + d * c.n_frequencies_dir);
    }
}

// ---------------- MLP ----------------
// y[o] = sum_i W[o][i] x[i]; accumulate fp32 (rounded to fp16 at the end) or per-16 fp16 blocks (wmma emulation)
inline float dot_acc(const float* w, const float* x, int n, int stride_w, bool accum16) {
    if (!accum16) {
        float acc = 0;
        for (int i = 0; i < n; i++) acc += w[(size_t)i * stride_w] * x[i];
        return h(acc);
    }
    float acc16 = 0;
    for (int i0 = 0; i0 < n; i0 += 16) {
        float blk = 0;
        for (int i = i0; i < std::min(n, i0 + 16); i++) blk += w[(size_t)i * stride_w] * x[i];
        acc16 = h(acc16 + blk);
    }
    return acc16;
}

// forward for one sample; acts (optional) receives [n_hidden][W] post-ReLU activations; out: [16]
void mlp_forward(const Model& m, const float* params16, const float* x, float* acts, float* out) {
    const int W = m.c.n_neurons, H = m.c.n_hidden_layers;
    const bool a16 = m.c.accum_fp16 != 0;
    std::vector<float> cur(x, x + m.in_w), nxt(W);
    for (int l = 0; l < H; l++) {
        const float* Wm = params16 + m.mat_off[l];
        const int cols = m.mat_cols[l];
        for (int o = 0; o < W; o++) {
            float v = dot_acc(Wm + (size_t)o * cols, cur.data(), cols, 1, a16);
            nxt[o] = v > 0 ? v : 0.0f;
        }
        if (acts) std::memcpy(acts + (size_t)l * W, nxt.data(), W * sizeof(float));
        cur.assign(nxt.begin(), nxt.end());
    }
    const float* Wo = params16 + m.mat_off[H];
    for (int o = 0; o < m.out_pad; o++) out[o] = dot_acc(Wo + (size_t)o * W, cur.data(), W, 1, a16);
}

}  // namespace

extern "C" {

void* nrco_create(const NrcoConfig* cfg, uint32_t seed) {
    Model* m = new Model();
    m->c = *cfg;
    derive(*m);
    const size_t P = m->n_params;
    m->master.assign(P, 0); m->w16.assign(P, 0); m->g16.assign(P, 0); m->ema16.assign(P, 0);   // ema memset 0 (ema.h:93-94)
    m->m1.assign(P, 0); m->m2.assign(P, 0); m->steps.assign(P, 0);
    init_params(*m, seed);
    return m;
}
void nrco_destroy(void* p) { delete (Model*)p; }
uint64_t nrco_n_params(void* p) { return ((Model*)p)->n_params; }
uint64_t nrco_n_mlp_params(void* p) { return ((Model*)p)->n_mlp; }
int32_t nrco_input_width(void* p) { return ((Model*)p)->in_w; }
uint32_t nrco_seed_word(uint32_t seed) { std::seed_seq s{seed}; std::vector<uint32_t> v(2); s.generate(v.begin(), v.end()); return v[0]; }
void nrco_grid_offsets(void* p, uint32_t* out) { Model* m = (Model*)p; for (size_t i = 0; i < m->offsets.size(); i++) out[i] = m->offsets[i]; }

// which: 0 master fp32, 1 working fp16, 2 ema fp16, 3 gradient fp16, 4 adam m, 5 adam v, 6 per-param step
void nrco_get(void* p, int which, float* out) {
    Model* m = (Model*)p;
    const std::vector<float>* src[6] = {&m->master, &m->w16, &m->ema16, &m->g16, &m->m1, &m->m2};
    if (which < 6) std::memcpy(out, src[which]->data(), m->n_params * sizeof(float));
    else for (size_t i = 0; i < m->n_params; i++) out[i] = (float)m->steps[i];
}
void nrco_set_params(void* p, const float* master) {      // Trainer::set_params_full_precision equivalent
    Model* m = (Model*)p;
    std::memcpy(m->master.data(), master, m->n_params * sizeof(float));
    for (size_t i = 0; i < m->n_params; i++) m->w16[i] = h(master[i]);
}
void nrco_set_ema(void* p, const float* ema) { Model* m = (Model*)p; for (size_t i = 0; i < m->n_params; i++) m->ema16[i] = h(ema[i]); }

void nrco_encode(void* p, const float* in, int n, int use_ema, float* out) {
    Model* m = (Model*)p;
    const float* params = use_ema ? m->ema16.data() : m->w16.data();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) encode(*m, params, in + 5 * (size_t)i, out + (size_t)m->in_w * i);
}

// NetworkWithInputEncoding::inference (object.h:147-176): EMA weights by default, fp16 net, float [n][3] out
void nrco_inference(void* p, const float* in, int n, int use_ema, float* out) {
    Model* m = (Model*)p;
    const float* params = use_ema ? m->ema16.data() : m->w16.data();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        std::vector<float> x(m->in_w); float o[16];
        encode(*m, params, in + 5 * (size_t)i, x.data());
        mlp_forward(*m, params, x.data(), nullptr, o);
        for (int k = 0; k < 3; k++) out[3 * (size_t)i + k] = o[k];
    }
}

// One Trainer::training_step (trainer.h:163-190) + Trainer::loss (trainer.h:205-207).
// run_optimizer==0 stops after backward (gradients inspectable through nrco_get(3)).
float nrco_training_step(void* p, const float* in, const float* target, int B, int run_optimizer) {
    Model* m = (Model*)p;
    const NrcoConfig& c = m->c;
    const int W = c.n_neurons, H = c.n_hidden_layers, IW = m->in_w, OP = m->out_pad;
    const float* wts = m->w16.data();
    std::vector<float> X((size_t)B * IW), A((size_t)B * H * W), O((size_t)B * OP), dO((size_t)B * OP, 0.0f);
    std::vector<double> Lvals(B, 0.0);
    // forward + loss (relative_l2_luminance.h:40-88); n_total = B*3
    const uint32_t n_total = (uint32_t)B * 3;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        encode(*m, wts, in + 5 * (size_t)b, X.data() + (size_t)b * IW);
        mlp_forward(*m, wts, X.data() + (size_t)b * IW, A.data() + (size_t)b * H * W, O.data() + (size_t)b * OP);
        const float* o = O.data() + (size_t)b * OP;
        const float lum = 0.299f * o[0] + 0.587f * o[1] + 0.114f * o[2];
        const float denom = lum * lum + 0.01f;
        double l = 0;
        for (int k = 0; k < 3; k++) {
            const float diff = o[k] - target[3 * (size_t)b + k];
            l += (double)(diff * diff / denom / 1.0f / (float)n_total);
            const float grad = 2 * diff / denom / 1.0f;
            dO[(size_t)b * OP + k] = h(c.loss_scale * grad / (float)n_total);
        }
        Lvals[b] = l;
    }
    double loss = 0; for (int b = 0; b < B; b++) loss += Lvals[b];

    // backward through the MLP (fully_fused_mlp.cu:150-259, 735-836)
    std::vector<float> dA((size_t)B * W), dAn((size_t)B * W);
    std::fill(m->g16.begin(), m->g16.end(), 0.0f);
    auto weight_grad = [&](int mat, const float* dY, int dy_stride, const float* Ain, int a_stride) {
        const int rows = m->mat_rows[mat], cols = m->mat_cols[mat];
        float* g = m->g16.data() + m->mat_off[mat];
#pragma omp parallel for schedule(static)
        for (int o = 0; o < rows; o++)
            for (int i = 0; i < cols; i++) {
                double acc = 0;
                for (int b = 0; b < B; b++) acc += (double)dY[(size_t)b * dy_stride + o] * (double)Ain[(size_t)b * a_stride + i];
                g[(size_t)o * cols + i] = hd(acc);
            }
    };
    const bool a16 = c.accum_fp16 != 0;
    // output layer
    weight_grad(H, dO.data(), OP, A.data() + (size_t)(H - 1) * W, H * W);
    {
        const float* Wo = wts + m->mat_off[H];
#pragma omp parallel for schedule(static)
        for (int b = 0; b < B; b++)
            for (int i = 0; i < W; i++) {
                float v = dot_acc(Wo + i, dO.data() + (size_t)b * OP, OP, W, a16);
                dA[(size_t)b * W + i] = A[(size_t)b * H * W + (size_t)(H - 1) * W + i] > 0 ? v : 0.0f;
            }
    }
    for (int l = H - 1; l >= 1; l--) {
        weight_grad(l, dA.data(), W, A.data() + (size_t)(l - 1) * W, H * W);
        const float* Wl = wts + m->mat_off[l];
#pragma omp parallel for schedule(static)
        for (int b = 0; b < B; b++)
            for (int i = 0; i < W; i++) {
                float v = dot_acc(Wl + i, dA.data() + (size_t)b * W, W, W, a16);
                dAn[(size_t)b * W + i] = A[(size_t)b * H * W + (size_t)(l - 1) * W + i] > 0 ? v : 0.0f;
            }
        dA.swap(dAn);
    }
    weight_grad(0, dA.data(), W, X.data(), IW);
    // dL/d(network input) = W0^T dA1 (fully_fused_mlp.cu:832-835), only needed for encoding params
    m->last_dL_dinput.assign((size_t)B * IW, 0.0f);
    if (m->n_grid) {
        const float* W0 = wts + m->mat_off[0];
#pragma omp parallel for schedule(static)
        for (int b = 0; b < B; b++)
            for (int i = 0; i < IW; i++) m->last_dL_dinput[(size_t)b * IW + i] = dot_acc(W0 + i, dA.data() + (size_t)b * W, W, IW, a16);
        // grid backward (grid.h:215-320): fp16 atomics emulated sequentially in sample order
        float* gg = m->g16.data();
        for (int b = 0; b < B; b++)
            for (int l = 0; l < c.n_levels; l++) {
                GridCell cell; grid_cell(*m, l, in + 5 * (size_t)b, &cell);
                const float g0 = m->last_dL_dinput[(size_t)b * IW + 2 * l], g1 = m->last_dL_dinput[(size_t)b * IW + 2 * l + 1];
                for (int k = 0; k < 8; k++) {
                    const float w = h(cell.w[k]);
                    float* dst = gg + m->n_mlp + cell.idx[k];
                    dst[0] = h(dst[0] + h(w * g0));
                    dst[1] = h(dst[1] + h(w * g1));
                }
            }
    }
    m->last_output = O; m->last_dL_doutput = dO;

    if (run_optimizer) {
        // Adam (adam.h:48-121) then EMA (ema.h:63-76, 102-138)
        m->current_step++;
        const float ema_debias_old = 1 - (float)std::pow(c.ema_decay, m->current_step - 1);
        const float ema_debias_new = 1.0f / (1 - (float)std::pow(c.ema_decay, m->current_step));
        const size_t P = m->n_params, NM = m->n_mlp;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < P; i++) {
            float gradient = m->g16[i] / c.loss_scale;
            bool skip = (i >= NM) && gradient == 0;
            if (!skip) {
                const float wfp = m->master[i];
                if (i < NM) gradient += c.l2_reg * wfp;
                const float gsq = gradient * gradient;
                const float fm = m->m1[i] = c.beta1 * m->m1[i] + (1 - c.beta1) * gradient;
                const float sm = m->m2[i] = c.beta2 * m->m2[i] + (1 - c.beta2) * gsq;
                float lr = c.learning_rate;
                const uint32_t st = ++m->steps[i];
                lr *= sqrtf(1 - powf(c.beta2, (float)st)) / (1 - powf(c.beta1, (float)st));
                const float eff = fminf(fmaxf(lr / (sqrtf(sm) + c.epsilon), 0.0f), 3.402823466e+38f);
                const float nw = wfp - eff * fm;
                m->master[i] = nw;
                m->w16[i] = h(nw);
            }
            const float filtered = (m->ema16[i] * c.ema_decay * ema_debias_old + m->w16[i] * (1 - c.ema_decay)) * ema_debias_new;
            m->ema16[i] = h(filtered);
        }
    }
    return (float)loss;
}

// inspection hooks for tests: last training step's padded output [B][16], dL/doutput [B][16], dL/dinput [B][in_w]
void nrco_last(void* p, int which, float* out) {
    Model* m = (Model*)p;
    const std::vector<float>& v = which == 0 ? m->last_output : which == 1 ? m->last_dL_doutput : m->last_dL_dinput;
    std::memcpy(out, v.data(), v.size() * sizeof(float));
}

float nrco_half_round(float x) { return h(x); }

}  // extern "C"
