// Plain-data types shared by the NRC host code and kernels (no device code in here).
#pragma once
#include <cstdint>

namespace nrchpm {

constexpr int kMaxLevels = 16;
constexpr int kWidth = 64;     // n_neurons
constexpr int kOutPad = 16;    // padded output width (fully_fused_mlp.cu:656)
constexpr int kTile = 128;     // records per warpgroup tile
constexpr int kFwdThreads = 256;

enum PosEnc { POS_HASHGRID = 0, POS_IDENTITY = 1, POS_TRIANGLE = 2, POS_FREQUENCY = 3 };
enum DirEnc { DIR_ONEBLOB = 0, DIR_IDENTITY = 1, DIR_TRIANGLE = 2 };

struct EncParams {
    int pos_enc, dir_enc;
    int pos_w, dir_w, in_w, dir_off;
    int n_levels, n_freq_pos, n_freq_dir, n_bins;
    int soa_bug;        // reproduce oneblob.h:224-227; only when oneblob_soa
    int oneblob_soa;    // SoA OneBlob kernel (composite.h:400-403: layout of the first nested encoding == HashGrid)
    float level_scale[kMaxLevels];
    uint32_t level_hsize[kMaxLevels];
    uint32_t level_offset[kMaxLevels];   // in grid entries (half2)
    uint32_t level_s0[kMaxLevels], level_s1[kMaxLevels], level_s2[kMaxLevels];   // dense strides with tcnn's uint32 wrap-around
    uint32_t level_hash[kMaxLevels];
    uint32_t level_kind[kMaxLevels];     // 1 dense, 2 hashed, 0 generic (uint32-wrapped strides: aliasing corners)
    int all_pow2;       // every level's table size is a power of two (`% size` is a mask)
};

// Adam state of ONE hash-grid entry (two fp16 features): fp32 master weights, first / second moments and the per-parameter
// step counters (adam.h:48-121 keeps four separate arrays).  32 bytes = one DRAM / L2 sector, so the optimizer touches one
// sector per updated entry instead of four.
struct alignas(32) GridAdamState {
    float master[2];
    float m1[2];
    float m2[2];
    uint32_t steps[2];
};

}  // namespace nrchpm
