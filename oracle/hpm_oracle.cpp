// TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle "A") of the reference's volumetric
// tracking passes.  Nothing under oracle/ is linked, imported or executed by the product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// The reference runs these passes as Vulkan GLSL compute shaders (there is no CUDA or CPU
// implementation to compile), so this file restates the GLSL line by line in scalar fp32 C++:
//   random.glsl:24-70        -> hash / floatConstruct / InitRandom / RandFloat
//   volume.glsl:1-39         -> sky_sdf / find_entry_exit / getDensity
//   dir_gen.glsl:1-64        -> hg_phase_func / rotationMatrix / NewRayDir
//   path_trace.glsl:24-174   -> RatioTrack / TraceDirLight / TracePointLight / SampleHdrEnvMap /
//                               TraceScene / DeltaTrack
//   nrc/gen_rays.comp:7-101, nrc/prep_infer_rays.comp:7-46, nrc/prep_train_rays.comp:7-138,
//   nrc/render.comp:7-41, nrc/clear.comp:5-9
// Parity status: "pinned" only through (a) the bit-exact RNG known answers in tests/ and (b) the
// statistical agreement of accumulated frames with the bundled reference/<scene>/0.exr statistics
// (tests/golden/exr_stats.json); GLSL transcendental rounding is driver-specific, so per-pixel
// bitwise parity with the reference's own frames is impossible by construction.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off: no FMA contraction, so + - * / sqrt are
// IEEE-exact and identical to the CUDA tracker compiled with -fmad=false).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

extern "C" {

struct HpmoScene {
    const uint8_t* grid;      // dense u8, index i + W*j + W*H*k  (Texture3D.cpp:107)
    int32_t dim[3];           // W,H,D
    float sky_size[3];        // normalize(extent)*107.5 (NrcHpmRenderer.cu:910-912)
    float density_factor;     // VOLUME_DENSITY_FACTOR
    float g;                  // VOLUME_G
    float dir_light_dir[3];
    float dir_light_strength;
    float point_pos[3];
    float point_strength;
    float point_color[3];
    float env_strength;       // HDR_ENV_MAP_STRENGTH
    float env_color[3];       // constant-colour env map (1x1 texture; black in the committed code, Q11)
};

struct HpmoConfig {
    uint32_t width, height;
    uint32_t train_width, train_height, train_x_dist, train_y_dist;
    uint32_t train_spp, primary_ray_length;
    float primary_ray_prob;
    uint32_t train_ring_size, train_ray_length;
    uint32_t infer_batch_size;
};

struct HpmoCamera {
    float inv_proj_view[16];  // column-major (glm)
    float pos[3];
};

}  // extern "C"

namespace {

constexpr float PI_F = 3.1415926535897932384626433832795028841971693993751058209749f;
constexpr float MAX_RAY_DISTANCE = 100000.0f;
constexpr float MIN_RAY_DISTANCE = 0.125f;

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length3(V3 a) { return std::sqrt(dot3(a, a)); }
inline V3 normalize3(V3 a) { float inv = 1.0f / std::sqrt(dot3(a, a)); return a * inv; }

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// random.glsl:24-33 -- one round of Bob Jenkins' one-at-a-time hash.
inline uint32_t hash1(uint32_t x) {
    x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
    return x;
}
inline uint32_t hash2(uint32_t a, uint32_t b) { return hash1(a ^ hash1(b)); }
inline uint32_t hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return hash1(a ^ hash1(b) ^ hash1(c) ^ hash1(d)); }
// random.glsl:42-52
inline float float_construct(uint32_t m) { return u2f((m & 0x007FFFFFu) | 0x3F800000u) - 1.0f; }

struct Ctx {
    const HpmoScene* sc;
    float rng;                // randomState (random.glsl:59)
    uint64_t lookups;         // density fetches (roofline accounting)
    V3 sky, half_sky, inv_sky;
    float inv_max_density;

    void init_random(float u, float v, const float fr[4]) {        // random.glsl:61-64
        float a = float_construct(hash2(f2u(u), f2u(v)));
        float b = float_construct(hash4(f2u(fr[0]), f2u(fr[1]), f2u(fr[2]), f2u(fr[3])));
        rng = float_construct(hash2(f2u(a), f2u(b)));
    }
    float rand_float(float max_val) {                                // random.glsl:66-70
        rng = float_construct(hash1(f2u(rng)));
        return rng * max_val;
    }
    float sky_sdf(V3 p) const {                                      // volume.glsl:1-5
        V3 d = {std::fabs(p.x) - half_sky.x, std::fabs(p.y) - half_sky.y, std::fabs(p.z) - half_sky.z};
        V3 m = {std::max(d.x, 0.0f), std::max(d.y, 0.0f), std::max(d.z, 0.0f)};
        return length3(m) + std::min(std::max(d.x, std::max(d.y, d.z)), 0.0f);
    }
    void find_entry_exit(V3 ro, V3 rd, V3* entry, V3* exit) const {   // volume.glsl:7-29
        float dist;
        do { dist = sky_sdf(ro); ro = ro + dist * rd; } while (dist > MIN_RAY_DISTANCE && dist < MAX_RAY_DISTANCE);
        *entry = ro;
        V3 two = sky * 2.0f;
        ro = ro + rd * length3(two);
        rd = rd * -1.0f;
        do { dist = sky_sdf(ro); ro = ro + dist * rd; } while (dist > MIN_RAY_DISTANCE && dist < MAX_RAY_DISTANCE);
        *exit = ro;
    }
    float get_density(V3 p) {                                        // volume.glsl:31-39 (+ Q9 sampler)
        lookups++;
        // p / skySize as p * (1 / skySize): what a GLSL compiler emits for a vector division, and what the CUDA path computes
        V3 uvw = V3{p.x * inv_sky.x + 0.5f, p.y * inv_sky.y + 0.5f, p.z * inv_sky.z + 0.5f};
        float fx = std::floor(uvw.x * (float)sc->dim[0]);
        float fy = std::floor(uvw.y * (float)sc->dim[1]);
        float fz = std::floor(uvw.z * (float)sc->dim[2]);
        float texel = 0.0f;                                          // clamp-to-border, opaque black
        if (fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < (float)sc->dim[0] && fy < (float)sc->dim[1] && fz < (float)sc->dim[2]) {
            size_t idx = (size_t)fx + (size_t)sc->dim[0] * ((size_t)fy + (size_t)sc->dim[1] * (size_t)fz);
            texel = (float)sc->grid[idx] / 255.0f;                   // UNORM8
        }
        return sc->density_factor * texel;
    }
    float hg_phase(float cos_theta) const {                          // dir_gen.glsl:1-7
        const float g = sc->g, g2 = g * g;
        return 0.5f * (1.0f - g2) / std::pow(1.0f + g2 - (2.0f * g * cos_theta), 1.5f);
    }
    // dir_gen.glsl:9-20 then `(rotMat * vec4(v, 1)).xyz`; GLSL mat4(...) is column-major.
    static V3 rotate(V3 axis, float angle, V3 v) {
        axis = normalize3(axis);
        float s = std::sin(angle), c = std::cos(angle), oc = 1.0f - c;
        V3 c0 = {oc * axis.x * axis.x + c, oc * axis.x * axis.y - axis.z * s, oc * axis.z * axis.x + axis.y * s};
        V3 c1 = {oc * axis.x * axis.y + axis.z * s, oc * axis.y * axis.y + c, oc * axis.y * axis.z - axis.x * s};
        V3 c2 = {oc * axis.z * axis.x - axis.y * s, oc * axis.y * axis.z + axis.x * s, oc * axis.z * axis.z + c};
        return (c0 * v.x + c1 * v.y) + c2 * v.z;
    }
    V3 new_ray_dir(V3 old_dir, bool phase_sampling) {                // dir_gen.glsl:22-64
        old_dir = normalize3(old_dir);
        V3 ortho = old_dir.z < old_dir.x ? V3{old_dir.y, -old_dir.x, 0.0f} : V3{0.0f, -old_dir.z, old_dir.y};
        // degenerate axis (old_dir exactly (-1,0,0) or (0,0,-1)): normalize(0) is undefined in GLSL; see DESIGN.md section 4
        if (ortho.x == 0.0f && ortho.y == 0.0f && ortho.z == 0.0f) ortho = V3{0.0f, 1.0f, 0.0f};
        ortho = normalize3(ortho);
        float angle;
        if (phase_sampling) {
            float g = sc->g, cos_theta;
            if (std::fabs(g) < 0.001f) {
                cos_theta = 1.0f - 2.0f * rand_float(1.0f);
            } else {
                float sqr_term = (1.0f - g * g) / (1.0f - g + (2.0f * g * rand_float(1.0f)));
                cos_theta = (1.0f + (g * g) - (sqr_term * sqr_term)) / (2.0f * g);
            }
            // GLSL leaves acos undefined outside [-1, 1]; fp32 rounding yields cos_theta = -1.0000004 for u == 0.  The bundled
            // reference/*/0.exr (8192 blended frames) contain no NaN, so the reference's driver returns a finite angle: clamp.
            angle = std::acos(std::fmin(1.0f, std::fmax(-1.0f, cos_theta)));
        } else {
            angle = rand_float(PI_F);
        }
        V3 nd = rotate(ortho, angle, old_dir);
        angle = rand_float(2.0f * PI_F);
        nd = rotate(old_dir, angle, nd);
        return normalize3(nd);
    }
    float ratio_track(V3 start, V3 end) {                            // path_trace.glsl:24-43
        V3 dir = normalize3(end - start);
        float t_max = length3(end - start);
        float transmittance = 1.0f, t = 0.0f;
        for (uint32_t i = 0; i < 128; i++) {
            t -= std::log(1.0f - rand_float(1.0f)) * inv_max_density;
            if (t >= t_max) break;
            V3 p = start + (t * dir);
            transmittance *= 1.0f - (get_density(p) * inv_max_density);
        }
        return transmittance;
    }
    V3 trace_dir_light(V3 pos, V3 dir) {                             // path_trace.glsl:45-56
        if (sc->dir_light_strength == 0.0f) return {0, 0, 0};
        V3 l = {sc->dir_light_dir[0], sc->dir_light_dir[1], sc->dir_light_dir[2]};
        V3 e, x;
        find_entry_exit(pos, neg(normalize3(l)), &e, &x);
        float tr = ratio_track(pos, x);
        float phase = hg_phase(dot3(l, neg(dir)));
        float v = 1.0f * tr * sc->dir_light_strength * phase;
        return {v, v, v};
    }
    V3 trace_point_light(V3 pos, V3 dir) {                           // path_trace.glsl:58-69
        if (sc->point_strength == 0.0f) return {0, 0, 0};
        V3 lp = {sc->point_pos[0], sc->point_pos[1], sc->point_pos[2]};
        float tr = ratio_track(lp, pos);
        float phase = hg_phase(dot3(normalize3(lp - pos), neg(dir)));
        V3 c = {sc->point_color[0], sc->point_color[1], sc->point_color[2]};
        return c * sc->point_strength * tr * phase;
    }
    V3 env_lookup() const {                                          // 1x1 texture * HDR_ENV_MAP_STRENGTH (path_trace.glsl:71-80)
        return V3{sc->env_color[0], sc->env_color[1], sc->env_color[2]} * sc->env_strength;
    }
    V3 sample_env(V3 pos, V3 dir) {                                  // path_trace.glsl:88-131, sampleCount == 1
        if (sc->env_strength == 0.0f) return {0, 0, 0};
        V3 rdir = new_ray_dir(dir, false);
        float phase = hg_phase(dot3(rdir, neg(dir)));
        V3 e, x;
        find_entry_exit(pos, rdir, &e, &x);
        float tr = ratio_track(pos, x);
        V3 light = env_lookup() * phase * tr;
        return light * (1.0f / 1.0f);
    }
    V3 trace_scene(V3 pos, V3 dir) {                                 // path_trace.glsl:133-137
        V3 a = trace_dir_light(pos, dir);
        V3 b = trace_point_light(pos, dir);
        V3 c = sample_env(pos, dir);
        return (a + b) + c;
    }
    V3 delta_track(V3 ro, V3 rd, bool* volume_exit) {                // path_trace.glsl:150-174
        *volume_exit = false;
        V3 e, x;
        find_entry_exit(ro, rd, &e, &x);
        float t_max = length3(x - ro);
        float t = 0.0f;
        for (uint32_t i = 0; i < 128; i++) {
            t -= std::log(1.0f - rand_float(1.0f)) * inv_max_density;
            if (t >= t_max) { *volume_exit = true; break; }
            V3 p = ro + (t * rd);
            if (get_density(p) * inv_max_density > rand_float(1.0f)) return p;
        }
        return ro + (rand_float(t_max) * rd);
    }
};

Ctx make_ctx(const HpmoScene* sc) {
    Ctx c{};
    c.sc = sc; c.rng = 0; c.lookups = 0;
    c.sky = {sc->sky_size[0], sc->sky_size[1], sc->sky_size[2]};
    c.half_sky = {sc->sky_size[0] / 2.0f, sc->sky_size[1] / 2.0f, sc->sky_size[2] / 2.0f};
    c.inv_sky = {1.0f / sc->sky_size[0], 1.0f / sc->sky_size[1], 1.0f / sc->sky_size[2]};
    c.inv_max_density = 1.0f / sc->density_factor;
    return c;
}

// prep_infer_rays.comp:7-24 / prep_train_rays.comp:38-54
void store_nrc_input(const Ctx& c, V3 pos, V3 dir, float* rec) {
    V3 np = pos / c.sky + c.sky * (1.0f / 2.0f);     // skySize / 2.0 is exact
    float theta = std::atan2(dir.z, dir.x);
    float norm_theta = (theta / PI_F) + 0.5f;
    float phi = std::acos(dir.y / std::sqrt(dir.x * dir.x + dir.z * dir.z));
    float norm_phi = phi / PI_F;
    rec[0] = np.x; rec[1] = np.y; rec[2] = np.z; rec[3] = norm_theta; rec[4] = norm_phi;
}

}  // namespace

extern "C" {

// --- known-answer hooks for the RNG (tests/test_oracle_rng.py) ---
uint32_t hpmo_hash(uint32_t x) { return hash1(x); }
float hpmo_float_construct(uint32_t m) { return float_construct(m); }
void hpmo_rng_stream(float u, float v, const float frame_random[4], int n, float* out) {
    HpmoScene sc{}; sc.density_factor = 1.0f; sc.sky_size[0] = sc.sky_size[1] = sc.sky_size[2] = 1.0f;
    Ctx c = make_ctx(&sc);
    c.init_random(u, v, frame_random);
    for (int i = 0; i < n; i++) out[i] = c.rand_float(1.0f);
}

// gen_rays.comp:73-80 only: per pixel, 1 when the primary ray reaches the volume (`!(sky_sdf(entry) > MAX_RAY_DISTANCE)` after the
// shader's own FindEntryExit march), 0 for a sky pixel.  Test aid: the CUDA tracker decides most sky pixels with a conservative
// analytic test instead of running the march, and tests/test_oracle_tracker.py checks that test against this function.
void hpmo_primary_hit(const HpmoScene* sc, const HpmoConfig* cfg, const HpmoCamera* cam, uint8_t* hit) {
    const uint32_t W = cfg->width, H = cfg->height;
    const float inv_w = 1.0f / (float)W, inv_h = 1.0f / (float)H;
    const float* M = cam->inv_proj_view;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t yy = 0; yy < (int64_t)H; yy++) {
        for (uint32_t x = 0; x < W; x++) {
            Ctx c = make_ctx(sc);
            const float u = (float)x * inv_w, v = (float)yy * inv_h;
            const float sx = (u * 2.0f) - 1.0f, sy = (v * 2.0f) - 1.0f, sz = 0.0f, sw = 1.0f;
            float wp[4];
            for (int r = 0; r < 4; r++) wp[r] = ((M[0 + r] * sx + M[4 + r] * sy) + M[8 + r] * sz) + M[12 + r] * sw;
            V3 pixel_world = {wp[0] / wp[3], wp[1] / wp[3], wp[2] / wp[3]};
            V3 ro = {cam->pos[0], cam->pos[1], cam->pos[2]};
            V3 rd = normalize3(pixel_world - ro);
            V3 entry, exit;
            c.find_entry_exit(ro, rd, &entry, &exit);
            hit[(size_t)yy * W + x] = c.sky_sdf(entry) > MAX_RAY_DISTANCE ? 0 : 1;
        }
    }
}

// gen_rays.comp main + TracePath.  Outputs are the four RGBA32F images of the reference, kept as
// planar-free AoS: primary_color[W*H][4] (rgb, factor), info[W*H] (didScatter as 0/1),
// nrc_origin[W*H][3], nrc_dir[W*H][3]; pixel (x,y) lives at y*W + x.  Returns #density lookups.
uint64_t hpmo_gen_rays(const HpmoScene* sc, const HpmoConfig* cfg, const HpmoCamera* cam, const float frame_random[4],
                       float* primary_color, float* info, float* nrc_origin, float* nrc_dir) {
    const uint32_t W = cfg->width, H = cfg->height;
    uint64_t total = 0;
    const float inv_w = 1.0f / (float)W, inv_h = 1.0f / (float)H;
    const float* M = cam->inv_proj_view;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
    for (int64_t yy = 0; yy < (int64_t)H; yy++) {
        for (uint32_t x = 0; x < W; x++) {
            const uint32_t y = (uint32_t)yy;
            Ctx c = make_ctx(sc);
            const float u = (float)x * inv_w, v = (float)y * inv_h;
            const float sx = (u * 2.0f) - 1.0f, sy = (v * 2.0f) - 1.0f, sz = 0.0f, sw = 1.0f;
            float wp[4];
            for (int r = 0; r < 4; r++) wp[r] = ((M[0 + r] * sx + M[4 + r] * sy) + M[8 + r] * sz) + M[12 + r] * sw;
            V3 pixel_world = {wp[0] / wp[3], wp[1] / wp[3], wp[2] / wp[3]};
            c.init_random(u, v, frame_random);
            V3 ro = {cam->pos[0], cam->pos[1], cam->pos[2]};
            V3 rd = normalize3(pixel_world - ro);
            V3 entry, exit;
            c.find_entry_exit(ro, rd, &entry, &exit);
            const size_t p = (size_t)y * W + x;
            float col[4];
            bool did_scatter = false;
            V3 env = c.env_lookup();
            if (c.sky_sdf(entry) > MAX_RAY_DISTANCE) {
                col[0] = env.x; col[1] = env.y; col[2] = env.z; col[3] = 1.0f;
            } else {
                // TracePath (gen_rays.comp:7-51)
                V3 light = {0, 0, 0};
                V3 e2, x2;
                c.find_entry_exit(ro, rd, &e2, &x2);
                V3 cur = e2, dir = rd;
                float factor = 1.0f;
                bool volume_exit = false;
                for (int i = 0; true; i++) {
                    cur = c.delta_track(cur, dir, &volume_exit);
                    if (volume_exit) break;
                    did_scatter = true;
                    factor *= 0.5f;
                    V3 l = c.trace_scene(cur, dir) * factor;
                    light = light + l;
                    dir = c.new_ray_dir(dir, true);
                    if (i >= (int)cfg->primary_ray_length) {
                        if (c.rand_float(1.0f) >= cfg->primary_ray_prob || i == 128) break;
                    }
                }
                nrc_origin[3 * p + 0] = cur.x; nrc_origin[3 * p + 1] = cur.y; nrc_origin[3 * p + 2] = cur.z;
                nrc_dir[3 * p + 0] = dir.x; nrc_dir[3 * p + 1] = dir.y; nrc_dir[3 * p + 2] = dir.z;
                col[0] = light.x; col[1] = light.y; col[2] = light.z; col[3] = factor;
                if (!did_scatter) { col[0] = env.x; col[1] = env.y; col[2] = env.z; col[3] = 1.0f; }
            }
            std::memcpy(primary_color + 4 * p, col, 16);
            info[p] = did_scatter ? 1.0f : 0.0f;
            total += c.lookups;
        }
    }
    return total;
}

// prep_infer_rays.comp main: records at index x*H + y, filter[idx / INFER_BATCH_SIZE] = 1.
// infer_input must be pre-zeroed by the caller (the reference vkCmdFillBuffer's it every frame).
void hpmo_prep_infer(const HpmoScene* sc, const HpmoConfig* cfg, const float* info, const float* nrc_origin,
                     const float* nrc_dir, float* infer_input, uint32_t* infer_filter) {
    Ctx c = make_ctx(sc);
    const uint32_t W = cfg->width, H = cfg->height;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x;
            if (info[p] != 1.0f) continue;
            const size_t lin = (size_t)x * H + y;
            V3 o = {nrc_origin[3 * p], nrc_origin[3 * p + 1], nrc_origin[3 * p + 2]};
            V3 d = {nrc_dir[3 * p], nrc_dir[3 * p + 1], nrc_dir[3 * p + 2]};
            store_nrc_input(c, o, d, infer_input + 5 * lin);
            infer_filter[lin / cfg->infer_batch_size] = 1;
        }
}

// prep_train_rays.comp main.  ring = {head, tail, RayInfo[train_w*train_h]} as uint32/float words.
// The reference's head/tail atomics race between invocations; this restatement fixes ONE legal
// schedule: all ring loads (in train-pixel order y*trainW+x) happen before all ring stores (same
// order).  The CUDA path implements the same schedule deterministically.
uint64_t hpmo_prep_train(const HpmoScene* sc, const HpmoConfig* cfg, const float frame_random[4], const float* info,
                         const float* nrc_origin, const float* nrc_dir, uint32_t* ring_words,
                         float* train_input, float* train_target) {
    const uint32_t W = cfg->width, H = cfg->height, TW = cfg->train_width, TH = cfg->train_height;
    const uint32_t ring_size = cfg->train_ring_size;
    uint32_t* head = ring_words + 0;
    uint32_t* tail = ring_words + 1;
    float* ring = reinterpret_cast<float*>(ring_words + 2);
    // clear.comp:5-9
    if (ring_size > 0) { *head %= ring_size; *tail %= ring_size; }
    const float inv_w = 1.0f / (float)W, inv_h = 1.0f / (float)H;
    uint64_t total = 0;
    const size_t T = (size_t)TW * TH;
    V3* org = new V3[T]; V3* dirs = new V3[T]; uint8_t* scat = new uint8_t[T];
    // phase 1: ray selection (pixel or ring load)
    for (uint32_t y = 0; y < TH; y++)
        for (uint32_t x = 0; x < TW; x++) {
            const size_t t = (size_t)y * TW + x;
            const uint32_t rx = x * cfg->train_x_dist, ry = y * cfg->train_y_dist;
            V3 o = {0, 0, 0};
            V3 d = normalize3(V3{1.0f, 1.0f, 1.0f});
            bool ds = false;
            if (rx < W && ry < H) ds = info[(size_t)ry * W + rx] == 1.0f;   // out-of-bounds imageLoad -> 0 (Q3)
            if (ds) {
                const size_t p = (size_t)ry * W + rx;
                o = {nrc_origin[3 * p], nrc_origin[3 * p + 1], nrc_origin[3 * p + 2]};
                d = {nrc_dir[3 * p], nrc_dir[3 * p + 1], nrc_dir[3 * p + 2]};
            } else if (ring_size > 0) {
                const uint32_t slot = ((*tail)++) % ring_size;
                o = {ring[6 * slot + 0], ring[6 * slot + 1], ring[6 * slot + 2]};
                d = {ring[6 * slot + 3], ring[6 * slot + 4], ring[6 * slot + 5]};
            }
            org[t] = o; dirs[t] = d; scat[t] = ds;
        }
    // phase 2: targets (independent per train pixel)
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : total)
    for (int64_t tt = 0; tt < (int64_t)T; tt++) {
        const uint32_t x = (uint32_t)(tt % TW), y = (uint32_t)(tt / TW);
        Ctx c = make_ctx(sc);
        c.init_random((float)x * inv_w, (float)y * inv_h, frame_random);
        V3 target = {0, 0, 0};
        for (uint32_t s = 0; s < cfg->train_spp; s++) {
            // TracePath (prep_train_rays.comp:69-99)
            V3 light = {0, 0, 0};
            V3 e, xx;
            c.find_entry_exit(org[tt], dirs[tt], &e, &xx);
            V3 cur = e, dir = dirs[tt];
            float factor = 1.0f;
            bool volume_exit = false;
            for (uint32_t i = 0; i < cfg->train_ray_length; i++) {
                cur = c.delta_track(cur, dir, &volume_exit);
                if (volume_exit) break;
                factor *= 0.5f;
                V3 l = c.trace_scene(cur, dir) * factor;
                light = light + l;
                dir = c.new_ray_dir(dir, true);
            }
            target = target + light;
        }
        { const float n = (float)cfg->train_spp; target = {target.x / n, target.y / n, target.z / n}; }   // prep_train_rays.comp:129
        if (ring_size > 0) {
            store_nrc_input(c, org[tt], dirs[tt], train_input + 5 * tt);
            train_target[3 * tt + 0] = std::min(8.0f, target.x);
            train_target[3 * tt + 1] = std::min(8.0f, target.y);
            train_target[3 * tt + 2] = std::min(8.0f, target.z);
        }
        total += c.lookups;
    }
    // phase 3: ring stores
    if (ring_size > 0)
        for (size_t t = 0; t < T; t++)
            if (scat[t]) {
                const uint32_t slot = ((*head)++) % ring_size;
                ring[6 * slot + 0] = org[t].x; ring[6 * slot + 1] = org[t].y; ring[6 * slot + 2] = org[t].z;
                ring[6 * slot + 3] = dirs[t].x; ring[6 * slot + 4] = dirs[t].y; ring[6 * slot + 5] = dirs[t].z;
            }
    delete[] org; delete[] dirs; delete[] scat;
    return total;
}

// render.comp main: output[W*H][4] is blended in place.
void hpmo_render(const HpmoConfig* cfg, const float* primary_color, const float* info, const float* infer_output,
                 uint32_t show_nrc, float blend_factor, float* output) {
    const uint32_t W = cfg->width, H = cfg->height;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            const size_t p = (size_t)y * W + x, lin = (size_t)x * H + y;
            float out[4] = {primary_color[4 * p], primary_color[4 * p + 1], primary_color[4 * p + 2], 1.0f};
            if (show_nrc == 1 && info[p] == 1.0f) {
                for (int k = 0; k < 3; k++) out[k] += std::max(0.0f, infer_output[3 * lin + k]) * primary_color[4 * p + 3];
            }
            for (int k = 0; k < 4; k++) output[4 * p + k] = (blend_factor * out[k]) + ((1.0f - blend_factor) * output[4 * p + k]);
        }
}

// mc/render.comp:7-84 restated for golden regeneration: plain path tracer,
// PATH_LENGTH vertices, rgb = scattered light (or env colour when nothing scattered), alpha = didScatter.
uint64_t hpmo_mc_render(const HpmoScene* sc, const HpmoConfig* cfg, const HpmoCamera* cam, const float frame_random[4],
                        uint32_t path_length, float blend_factor, float* output) {
    const uint32_t W = cfg->width, H = cfg->height;
    uint64_t total = 0;
    const float inv_w = 1.0f / (float)W, inv_h = 1.0f / (float)H;
    const float* M = cam->inv_proj_view;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
    for (int64_t yy = 0; yy < (int64_t)H; yy++) {
        for (uint32_t x = 0; x < W; x++) {
            const uint32_t y = (uint32_t)yy;
            Ctx c = make_ctx(sc);
            const float u = (float)x * inv_w, v = (float)y * inv_h;
            const float sx = (u * 2.0f) - 1.0f, sy = (v * 2.0f) - 1.0f;
            float wp[4];
            for (int r = 0; r < 4; r++) wp[r] = ((M[0 + r] * sx + M[4 + r] * sy) + M[8 + r] * 0.0f) + M[12 + r] * 1.0f;
            V3 pixel_world = {wp[0] / wp[3], wp[1] / wp[3], wp[2] / wp[3]};
            c.init_random(u, v, frame_random);
            V3 ro = {cam->pos[0], cam->pos[1], cam->pos[2]};
            V3 rd = normalize3(pixel_world - ro);
            V3 entry, exit;
            c.find_entry_exit(ro, rd, &entry, &exit);
            float col[4];
            V3 env = c.env_lookup();
            col[0] = env.x; col[1] = env.y; col[2] = env.z; col[3] = 0.0f;
            if (!(c.sky_sdf(entry) > MAX_RAY_DISTANCE)) {
                V3 light = {0, 0, 0};
                V3 cur = entry, dir = rd;
                float factor = 1.0f;
                bool volume_exit = false, did = false;
                for (uint32_t i = 0; i < path_length; i++) {
                    cur = c.delta_track(cur, dir, &volume_exit);
                    if (volume_exit) break;
                    did = true;
                    factor *= 0.5f;
                    light = light + c.trace_scene(cur, dir) * factor;
                    dir = c.new_ray_dir(dir, true);
                }
                if (did) { col[0] = light.x; col[1] = light.y; col[2] = light.z; col[3] = 1.0f; }
            }
            const size_t p = (size_t)y * W + x;
            for (int k = 0; k < 4; k++) output[4 * p + k] = (blend_factor * col[k]) + ((1.0f - blend_factor) * output[4 * p + k]);
            total += c.lookups;
        }
    }
    return total;
}

}  // extern "C"
