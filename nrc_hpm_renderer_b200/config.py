"""Host-side mirror of the reference's ``en::AppConfig`` (reference include/engine/AppConfig.hpp:10-66,
src/AppConfig.cpp:9-182): the 17 positional CLI arguments, the encoding presets (posID / dirID ->
tiny-cuda-nn JSON) and the scene presets, plus the derived renderer constants
(``NrcHpmRenderer::CalcTrainSubset``, reference src/NrcHpmRenderer.cu:612-642).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

# reference src/main.cu:432-439
DEFAULT_ARGV = ["NRC-HPM-Renderer", "RelativeL2Luminance", "Adam", "0.01", "0.99", "0", "0", "64", "6", "21", "14", "4",
                "4", "1.0", "1", "1", "0.0", "32"]


def encoding_json(pos_id: int, dir_id: int) -> dict:
    """AppConfig::NNEncodingConfig (reference src/AppConfig.cpp:9-80)."""
    pos = {
        0: {"otype": "HashGrid", "n_dims_to_encode": 3, "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19,
            "base_resolution": 16, "per_level_scale": 2.0},
        1: {"otype": "Identity", "n_dims_to_encode": 3},
        2: {"otype": "TriangleWave", "n_dims_to_encode": 3, "n_frequencies": 12},
        3: {"otype": "Frequency", "n_dims_to_encode": 3, "n_frequencies": 12},
    }
    dr = {
        0: {"otype": "OneBlob", "n_dims_to_encode": 2, "n_bins": 4},
        1: {"otype": "Identity", "n_dims_to_encode": 2},
        2: {"otype": "TriangleWave", "n_dims_to_encode": 2, "n_frequencies": 4},
    }
    if pos_id not in pos:
        raise RuntimeError("NNEncodingConfig posID is invalid")      # src/AppConfig.cpp:45
    if dir_id not in dr:
        raise RuntimeError("NNEncodingConfig dirID is invalid")      # src/AppConfig.cpp:71
    return {"otype": "Composite", "reduction": "Concatenation", "nested": [pos[pos_id], dr[dir_id]]}


@dataclass
class HpmSceneConfig:
    """AppConfig::HpmSceneConfig (reference src/AppConfig.cpp:86-150)."""
    id: int = 0
    dir_light_strength: float = 0.0
    point_light_strength: float = 0.0
    hdr_env_map_path: str = ""
    hdr_env_map_strength: float = 0.0
    density: float = 0.0
    dynamic: bool = False

    @staticmethod
    def preset(scene_id: int) -> "HpmSceneConfig":
        table = {  # id: (dir, point, env, density)
            0: (16.0, 0.0, 0.0, 0.6), 1: (0.0, 64.0, 0.0, 0.6), 2: (0.0, 128.0, 0.0, 1.0),
            3: (16.0, 0.0, 0.0, 0.25), 4: (8.0, 0.0, 0.1, 0.6), 5: (0.0, 0.0, 1.0, 1.6),
        }
        if scene_id not in table:
            raise RuntimeError("HpmSceneConfig ID is invalid")       # src/AppConfig.cpp:147
        d, p, e, rho = table[scene_id]
        return HpmSceneConfig(scene_id, d, p, "", e, rho, False)


@dataclass
class AppConfig:
    loss_fn: str = "RelativeL2Luminance"
    optimizer: str = "Adam"
    learning_rate: float = 0.01
    ema_decay: float = 0.99
    pos_enc_id: int = 0
    dir_enc_id: int = 0
    nn_width: int = 64
    nn_depth: int = 6
    log2_infer_batch_size: int = 21
    log2_train_batch_size: int = 14
    train_batch_count: int = 4
    scene: HpmSceneConfig = field(default_factory=lambda: HpmSceneConfig.preset(4))
    train_ring_buf_size: float = 1.0
    train_spp: int = 1
    primary_ray_length: int = 1
    primary_ray_prob: float = 0.0
    train_ray_length: int = 32

    @staticmethod
    def from_argv(argv) -> "AppConfig":
        """AppConfig(const std::vector<char*>&) -- exactly 18 entries (reference src/AppConfig.cpp:154-182)."""
        if len(argv) != 18:
            raise RuntimeError("Argument count does not match requirements for AppConfig")
        a = list(argv)[1:]
        enc_pos, enc_dir = int(a[4]), int(a[5])
        encoding_json(enc_pos, enc_dir)
        return AppConfig(str(a[0]), str(a[1]), float(a[2]), float(a[3]), enc_pos, enc_dir, int(a[6]), int(a[7]), int(a[8]),
                         int(a[9]), int(a[10]), HpmSceneConfig.preset(int(a[11])), float(a[12]), int(a[13]), int(a[14]),
                         float(a[15]), int(a[16]))

    @staticmethod
    def default() -> "AppConfig":
        return AppConfig.from_argv(DEFAULT_ARGV)

    # ---- reference src/NeuralRadianceCache.cu:12-13
    @property
    def infer_batch_size(self) -> int:
        return 2 << (self.log2_infer_batch_size - 1)

    @property
    def train_batch_size(self) -> int:
        return 2 << (self.log2_train_batch_size - 1)

    def model_json(self) -> dict:
        """The tiny-cuda-nn model JSON the reference builds (src/NeuralRadianceCache.cu:16-37)."""
        return {
            "loss": {"otype": self.loss_fn},
            "optimizer": {"otype": "EMA", "decay": self.ema_decay,
                          "nested": {"otype": self.optimizer, "learning_rate": self.learning_rate}},
            "encoding": encoding_json(self.pos_enc_id, self.dir_enc_id),
            "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None",
                        "n_neurons": self.nn_width, "n_hidden_layers": self.nn_depth},
        }

    def get_name(self) -> str:
        """AppConfig::GetName (src/AppConfig.cpp:184-205); std::to_string(float) prints 6 decimals."""
        f = lambda v: f"{v:.6f}"
        return "_".join([self.loss_fn, self.optimizer, f(self.learning_rate), f(self.ema_decay), str(self.pos_enc_id),
                         str(self.dir_enc_id), str(self.nn_width), str(self.nn_depth), str(self.log2_infer_batch_size),
                         str(self.log2_train_batch_size), str(self.train_batch_count), str(self.scene.id),
                         f(self.train_ring_buf_size), str(self.train_spp), str(self.primary_ray_length),
                         f(self.primary_ray_prob), str(self.train_ray_length)])


@dataclass
class TrainSubset:
    train_width: int
    train_height: int
    x_dist: int
    y_dist: int


def calc_train_subset(render_width: int, render_height: int, train_pixel_count: int) -> TrainSubset:
    """NrcHpmRenderer::CalcTrainSubset (reference src/NrcHpmRenderer.cu:612-642)."""
    root = int(math.sqrt(train_pixel_count))
    for factor in range(root, 1, -1):
        if train_pixel_count % factor == 0:
            other = train_pixel_count // factor
            big, small = max(factor, other), min(factor, other)
            tw, th = (big, small) if render_width > render_height else (small, big)
            return TrainSubset(tw, th, render_width // tw, render_height // th)
    raise RuntimeError("Could not find suitable division of trainPixelCount")


def sky_size(extent) -> tuple:
    """normalize(extent) * 107.5 in fp32 (reference src/NrcHpmRenderer.cu:910-912)."""
    import numpy as np
    e = np.asarray(extent, dtype=np.float32)
    inv = np.float32(1.0) / np.sqrt((e * e).sum(dtype=np.float32))
    return tuple(float(v) for v in (e * inv * np.float32(107.5)).astype(np.float32))
