"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/hpm_oracle.cpp,
oracle/nrc_oracle.cpp).  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; never by the product package ``nrc_hpm_renderer_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("hpm_oracle.cpp", "nrc_oracle.cpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


class HpmoScene(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("dim", C.c_int32 * 3), ("sky_size", C.c_float * 3), ("density_factor", C.c_float),
                ("g", C.c_float), ("dir_light_dir", C.c_float * 3), ("dir_light_strength", C.c_float),
                ("point_pos", C.c_float * 3), ("point_strength", C.c_float), ("point_color", C.c_float * 3),
                ("env_strength", C.c_float), ("env_color", C.c_float * 3)]


class HpmoConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("train_width", C.c_uint32), ("train_height", C.c_uint32),
                ("train_x_dist", C.c_uint32), ("train_y_dist", C.c_uint32), ("train_spp", C.c_uint32),
                ("primary_ray_length", C.c_uint32), ("primary_ray_prob", C.c_float), ("train_ring_size", C.c_uint32),
                ("train_ray_length", C.c_uint32), ("infer_batch_size", C.c_uint32)]


class HpmoCamera(C.Structure):
    _fields_ = [("inv_proj_view", C.c_float * 16), ("pos", C.c_float * 3)]


class NrcoConfig(C.Structure):
    _fields_ = [("pos_enc", C.c_int32), ("dir_enc", C.c_int32), ("n_neurons", C.c_int32), ("n_hidden_layers", C.c_int32),
                ("oneblob_soa_bug", C.c_int32), ("accum_fp16", C.c_int32), ("n_levels", C.c_int32),
                ("log2_hashmap_size", C.c_int32), ("base_resolution", C.c_int32), ("per_level_scale", C.c_float),
                ("n_frequencies_pos", C.c_int32), ("n_frequencies_dir", C.c_int32), ("n_bins", C.c_int32),
                ("learning_rate", C.c_float), ("ema_decay", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("epsilon", C.c_float), ("l2_reg", C.c_float), ("loss_scale", C.c_float)]


def _declare(l):
    fp = C.POINTER(C.c_float)
    l.hpmo_hash.restype = C.c_uint32; l.hpmo_hash.argtypes = [C.c_uint32]
    l.hpmo_float_construct.restype = C.c_float; l.hpmo_float_construct.argtypes = [C.c_uint32]
    l.hpmo_rng_stream.argtypes = [C.c_float, C.c_float, fp, C.c_int, fp]
    l.hpmo_gen_rays.restype = C.c_uint64
    l.hpmo_gen_rays.argtypes = [C.POINTER(HpmoScene), C.POINTER(HpmoConfig), C.POINTER(HpmoCamera), fp, fp, fp, fp, fp]
    l.hpmo_primary_hit.argtypes = [C.POINTER(HpmoScene), C.POINTER(HpmoConfig), C.POINTER(HpmoCamera), C.POINTER(C.c_uint8)]
    l.hpmo_prep_infer.argtypes = [C.POINTER(HpmoScene), C.POINTER(HpmoConfig), fp, fp, fp, fp, C.POINTER(C.c_uint32)]
    l.hpmo_prep_train.restype = C.c_uint64
    l.hpmo_prep_train.argtypes = [C.POINTER(HpmoScene), C.POINTER(HpmoConfig), fp, fp, fp, fp, C.POINTER(C.c_uint32), fp, fp]
    l.hpmo_render.argtypes = [C.POINTER(HpmoConfig), fp, fp, fp, C.c_uint32, C.c_float, fp]
    l.hpmo_mc_render.restype = C.c_uint64
    l.hpmo_mc_render.argtypes = [C.POINTER(HpmoScene), C.POINTER(HpmoConfig), C.POINTER(HpmoCamera), fp, C.c_uint32, C.c_float, fp]
    l.nrco_create.restype = C.c_void_p; l.nrco_create.argtypes = [C.POINTER(NrcoConfig), C.c_uint32]
    l.nrco_destroy.argtypes = [C.c_void_p]
    l.nrco_n_params.restype = C.c_uint64; l.nrco_n_params.argtypes = [C.c_void_p]
    l.nrco_n_mlp_params.restype = C.c_uint64; l.nrco_n_mlp_params.argtypes = [C.c_void_p]
    l.nrco_input_width.restype = C.c_int32; l.nrco_input_width.argtypes = [C.c_void_p]
    l.nrco_seed_word.restype = C.c_uint32; l.nrco_seed_word.argtypes = [C.c_uint32]
    l.nrco_grid_offsets.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    l.nrco_get.argtypes = [C.c_void_p, C.c_int, fp]
    l.nrco_set_params.argtypes = [C.c_void_p, fp]
    l.nrco_set_ema.argtypes = [C.c_void_p, fp]
    l.nrco_encode.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, fp]
    l.nrco_inference.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, fp]
    l.nrco_training_step.restype = C.c_float; l.nrco_training_step.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_int]
    l.nrco_last.argtypes = [C.c_void_p, C.c_int, fp]
    l.nrco_half_round.restype = C.c_float; l.nrco_half_round.argtypes = [C.c_float]


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


# ------------------------------------------------------------------ tracker
def make_scene(grid_u8: np.ndarray, sky_size, density, g=0.8, dir_light_dir=(0.0, 0.0, -1.0), dir_light_strength=0.0,
               point_pos=(0.0, 0.0, 0.0), point_strength=0.0, point_color=(1.0, 1.0, 1.0), env_strength=0.0,
               env_color=(0.0, 0.0, 0.0)) -> HpmoScene:
    assert grid_u8.dtype == np.uint8 and grid_u8.flags["C_CONTIGUOUS"] and grid_u8.ndim == 3
    d, h, w = grid_u8.shape
    s = HpmoScene()
    s.grid = grid_u8.ctypes.data
    s._keepalive = grid_u8
    s.dim[:] = (w, h, d)
    s.sky_size[:] = [float(v) for v in sky_size]
    s.density_factor, s.g = float(density), float(g)
    s.dir_light_dir[:] = [float(v) for v in dir_light_dir]; s.dir_light_strength = float(dir_light_strength)
    s.point_pos[:] = [float(v) for v in point_pos]; s.point_strength = float(point_strength)
    s.point_color[:] = [float(v) for v in point_color]
    s.env_strength = float(env_strength); s.env_color[:] = [float(v) for v in env_color]
    return s


def make_config(width, height, train_width, train_height, train_x_dist, train_y_dist, train_spp=1, primary_ray_length=1,
                primary_ray_prob=0.0, train_ring_size=0, train_ray_length=1, infer_batch_size=1 << 21) -> HpmoConfig:
    return HpmoConfig(width, height, train_width, train_height, train_x_dist, train_y_dist, train_spp, primary_ray_length,
                      float(primary_ray_prob), train_ring_size, train_ray_length, infer_batch_size)


def make_camera(inv_proj_view_colmajor: np.ndarray, pos) -> HpmoCamera:
    c = HpmoCamera()
    c.inv_proj_view[:] = [float(v) for v in np.asarray(inv_proj_view_colmajor, dtype=np.float32).reshape(16)]
    c.pos[:] = [float(v) for v in pos]
    return c


def rng_stream(u, v, frame_random, n):
    out = np.empty(n, dtype=np.float32)
    fr = np.asarray(frame_random, dtype=np.float32)
    lib().hpmo_rng_stream(C.c_float(u), C.c_float(v), _fp(fr), n, _fp(out))
    return out


def gen_rays(scene, cfg, cam, frame_random):
    n = cfg.width * cfg.height
    color = np.zeros((n, 4), np.float32); info = np.zeros(n, np.float32)
    org = np.zeros((n, 3), np.float32); dr = np.zeros((n, 3), np.float32)
    fr = np.asarray(frame_random, dtype=np.float32)
    lookups = lib().hpmo_gen_rays(C.byref(scene), C.byref(cfg), C.byref(cam), _fp(fr), _fp(color), _fp(info), _fp(org), _fp(dr))
    return dict(color=color, info=info, origin=org, dir=dr, lookups=int(lookups))


def primary_hit(scene, cfg, cam) -> np.ndarray:
    """per pixel [H][W]: 1 when the shader's FindEntryExit march reaches the volume, 0 for a sky pixel (gen_rays.comp:73-80)"""
    hit = np.zeros((cfg.height, cfg.width), np.uint8)
    lib().hpmo_primary_hit(C.byref(scene), C.byref(cfg), C.byref(cam), hit.ctypes.data_as(C.POINTER(C.c_uint8)))
    return hit


def prep_infer(scene, cfg, rays):
    n = cfg.width * cfg.height
    rec = np.zeros((n, 5), np.float32)
    filt = np.zeros((n + cfg.infer_batch_size - 1) // cfg.infer_batch_size, np.uint32)
    lib().hpmo_prep_infer(C.byref(scene), C.byref(cfg), _fp(rays["info"]), _fp(rays["origin"]), _fp(rays["dir"]), _fp(rec),
                          filt.ctypes.data_as(C.POINTER(C.c_uint32)))
    return rec, filt


def new_ring(cfg) -> np.ndarray:
    """NrcHpmRenderer::CreateNrcTrainRingBuffer (reference src/NrcHpmRenderer.cu:841-881): head, tail, then
    train_w*train_h RayInfo initialised to pos 0, dir (0,0,1)."""
    t = cfg.train_width * cfg.train_height
    words = np.zeros(2 + 6 * t, np.uint32)
    rays = words[2:].view(np.float32).reshape(t, 6)
    rays[:, 5] = 1.0
    return words


def prep_train(scene, cfg, rays, frame_random, ring_words):
    t = cfg.train_width * cfg.train_height
    tin = np.zeros((t, 5), np.float32); tgt = np.zeros((t, 3), np.float32)
    fr = np.asarray(frame_random, dtype=np.float32)
    lookups = lib().hpmo_prep_train(C.byref(scene), C.byref(cfg), _fp(fr), _fp(rays["info"]), _fp(rays["origin"]), _fp(rays["dir"]),
                                    ring_words.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(tin), _fp(tgt))
    return tin, tgt, int(lookups)


def render(cfg, rays, infer_output, show_nrc, blend_factor, output):
    lib().hpmo_render(C.byref(cfg), _fp(rays["color"]), _fp(rays["info"]), _fp(infer_output), int(show_nrc), C.c_float(blend_factor), _fp(output))
    return output


def mc_render(scene, cfg, cam, frame_random, path_length, blend_factor, output):
    fr = np.asarray(frame_random, dtype=np.float32)
    return int(lib().hpmo_mc_render(C.byref(scene), C.byref(cfg), C.byref(cam), _fp(fr), int(path_length), C.c_float(blend_factor), _fp(output)))


# ------------------------------------------------------------------ NRC
def compare_images(ref_rgba: np.ndarray, cmp_rgba: np.ndarray) -> dict:
    """Reference::Result restated in numpy (float64 sums): data/shader/ref/cmp1.comp:25-41 (mse, refMean, ownMean, count over the
    pixels with ref alpha != 0), norm.comp:19-24 (divide by the count), cmp2.comp:25-40 (ownVar around the scalar ownMean)."""
    ref, cmp_ = np.asarray(ref_rgba, np.float32).reshape(-1, 4), np.asarray(cmp_rgba, np.float32).reshape(-1, 4)
    valid = ref[:, 3] != 0.0
    n = int(valid.sum())
    r, c = ref[valid, :3].astype(np.float64), cmp_[valid, :3].astype(np.float64)
    if n == 0:
        return {"mse": 0.0, "refMean": 0.0, "ownMean": 0.0, "ownVar": 0.0, "validPixelCount": 0}
    own_mean = float((c.sum(1) / 3).sum() / n)
    return {"mse": float((((c - r) ** 2).sum(1) / 3).sum() / n), "refMean": float((r.sum(1) / 3).sum() / n), "ownMean": own_mean,
            "ownVar": float((((c - np.float64(np.float32(own_mean))) ** 2).sum(1) / 3).sum() / n), "validPixelCount": n}


def nrc_config(pos_enc=0, dir_enc=0, n_hidden_layers=6, n_neurons=64, oneblob_soa_bug=1, accum_fp16=0, learning_rate=0.01,
               ema_decay=0.99) -> NrcoConfig:
    return NrcoConfig(pos_enc, dir_enc, n_neurons, n_hidden_layers, oneblob_soa_bug, accum_fp16, 16, 19, 16, 2.0, 12, 4, 4,
                      learning_rate, ema_decay, 0.9, 0.999, 1e-8, 1e-8, 128.0)


class NrcOracle:
    """Stateful CPU restatement of tcnn's TrainableModel as the reference drives it."""
    MASTER, WORKING, EMA, GRAD, ADAM_M, ADAM_V, STEPS = range(7)

    def __init__(self, cfg: NrcoConfig, seed: int = 1337):
        self.cfg = cfg
        self._h = lib().nrco_create(C.byref(cfg), seed)
        self.n_params = int(lib().nrco_n_params(self._h))
        self.n_mlp = int(lib().nrco_n_mlp_params(self._h))
        self.input_width = int(lib().nrco_input_width(self._h))

    def __del__(self):
        try:
            lib().nrco_destroy(self._h)
        except Exception:
            pass

    def get(self, which) -> np.ndarray:
        out = np.empty(self.n_params, np.float32)
        lib().nrco_get(self._h, which, _fp(out))
        return out

    def grid_offsets(self):
        out = np.zeros(self.cfg.n_levels + 1, np.uint32)
        lib().nrco_grid_offsets(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def set_params(self, master: np.ndarray):
        lib().nrco_set_params(self._h, _fp(np.ascontiguousarray(master, np.float32)))

    def set_ema(self, ema: np.ndarray):
        lib().nrco_set_ema(self._h, _fp(np.ascontiguousarray(ema, np.float32)))

    def encode(self, rec: np.ndarray, use_ema=False) -> np.ndarray:
        rec = np.ascontiguousarray(rec, np.float32)
        out = np.empty((len(rec), self.input_width), np.float32)
        lib().nrco_encode(self._h, _fp(rec), len(rec), int(use_ema), _fp(out))
        return out

    def inference(self, rec: np.ndarray, use_ema=True) -> np.ndarray:
        rec = np.ascontiguousarray(rec, np.float32)
        out = np.empty((len(rec), 3), np.float32)
        lib().nrco_inference(self._h, _fp(rec), len(rec), int(use_ema), _fp(out))
        return out

    def training_step(self, rec: np.ndarray, target: np.ndarray, run_optimizer=True) -> float:
        rec = np.ascontiguousarray(rec, np.float32); target = np.ascontiguousarray(target, np.float32)
        return float(lib().nrco_training_step(self._h, _fp(rec), _fp(target), len(rec), int(run_optimizer)))

    def last(self, which, batch):
        width = 16 if which < 2 else self.input_width
        out = np.empty((batch, width), np.float32)
        lib().nrco_last(self._h, which, _fp(out))
        return out
