"""Python mirrors of ``en::HpmScene`` (reference src/HpmScene.cpp:23-54), ``en::NrcHpmRenderer`` (reference
include/engine/graphics/renderer/NrcHpmRenderer.hpp:13-41, src/NrcHpmRenderer.cu:212-353) and ``en::McHpmRenderer``
(src/McHpmRenderer.cpp:121-151) on top of the C ABI.  ``VkQueue`` / ``VkDevice`` arguments of the reference have no
counterpart: the passes run on one CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from .camera import Camera
from .config import AppConfig, HpmSceneConfig, calc_train_subset, sky_size
from .nrc import NeuralRadianceCache

(BUF_OUTPUT, BUF_PRIMARY_COLOR, BUF_PRIMARY_INFO, BUF_NRC_ORIGIN, BUF_NRC_DIR, BUF_INFER_INPUT, BUF_INFER_OUTPUT,
 BUF_TRAIN_INPUT, BUF_TRAIN_TARGET, BUF_TRAIN_RING, BUF_INFER_FILTER, BUF_COUNTERS) = range(12)
STAGE_NAMES = ("clear", "gen_rays", "prep_train_rays", "nrc_inference", "nrc_train", "render", "total")


def dir_light_vec(zenith: float, azimuth: float):
    """DirLight::VecFromAngles (reference src/DirLight.cpp:5-14): rotY(azimuth) * rotX(zenith) * (0,1,0), glm fp32."""
    z, a = np.float32(zenith), np.float32(azimuth)
    # rotate (0,1,0) about X by zenith: (0, cos z, sin z); then about Y by azimuth
    v = np.array([0.0, np.cos(z), np.sin(z)], dtype=np.float32)
    ca, sa = np.cos(a), np.sin(a)
    return (float(ca * v[0] + sa * v[2]), float(v[1]), float(-sa * v[0] + ca * v[2]))


class HpmScene:
    """Scene container: density grid + the light / medium parameters of a reference scene preset."""

    def __init__(self, grid_u8: np.ndarray, scene: HpmSceneConfig | int = 0, *, g: float = 0.8, env_color=(0.0, 0.0, 0.0),
                 dir_light_dir=None, point_pos=(0.0, 0.0, 0.0), point_color=(1.0, 1.0, 1.0), density: float | None = None):
        assert grid_u8.dtype == np.uint8 and grid_u8.ndim == 3 and grid_u8.flags["C_CONTIGUOUS"]
        if isinstance(scene, int):
            scene = HpmSceneConfig.preset(scene)
        d, h, w = grid_u8.shape
        self.dims = (w, h, d)
        self.sky_size = sky_size((w, h, d))
        self.config = scene
        desc = _lib.SceneDesc()
        desc.dim[:] = (w, h, d)
        desc.sky_size[:] = self.sky_size
        desc.density_factor = float(scene.density if density is None else density)
        desc.g = float(g)
        # src/HpmScene.cpp:28: DirLight(-1.57, 0.0, ...) ; committed code
        dl = dir_light_vec(-1.57, 0.0) if dir_light_dir is None else dir_light_dir
        desc.dir_light_dir[:] = [float(v) for v in dl]
        desc.dir_light_strength = float(scene.dir_light_strength)
        desc.point_pos[:] = [float(v) for v in point_pos]
        desc.point_strength = float(scene.point_light_strength)
        desc.point_color[:] = [float(v) for v in point_color]
        desc.env_strength = float(scene.hdr_env_map_strength)
        desc.env_color[:] = [float(v) for v in env_color]
        self.desc = desc
        self._h = C.c_void_p()
        _lib.check(_lib.lib().hpm_scene_create(C.byref(desc), grid_u8.ctypes.data_as(C.c_void_p), C.byref(self._h)))

    def Destroy(self):
        if self._h:
            _lib.check(_lib.lib().hpm_scene_destroy(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass


def make_render_config(width, height, app: AppConfig, *, blend=False, show_nrc=True, compact_inference=True, train_pixels=None,
                       parity_q2=True, parity_q3=True, x_begin=0, x_end=0, pipeline_train=False) -> _lib.RenderConfig:
    """Specialization constants of NrcHpmRenderer::InitSpecializationConstants (reference src/NrcHpmRenderer.cu:908-1061).
    parity_q2: the reference never passes TRAIN_RAY_LENGTH, shaders see 1.
    parity_q3: TRAIN_Y_DIST receives trainXDist."""
    t = app.train_batch_count * app.train_batch_size if train_pixels is None else train_pixels
    c = _lib.RenderConfig()
    c.width, c.height = width, height
    if t:
        ts = calc_train_subset(width, height, t)
        c.train_width, c.train_height = ts.train_width, ts.train_height
        c.train_x_dist = ts.x_dist
        c.train_y_dist = ts.x_dist if parity_q3 else ts.y_dist
        c.train_ring_size = int(app.train_ring_buf_size * float(ts.train_width * ts.train_height))   # :251
    c.train_spp = app.train_spp
    c.primary_ray_length = app.primary_ray_length
    c.primary_ray_prob = app.primary_ray_prob
    c.train_ray_length = 1 if parity_q2 else app.train_ray_length
    c.infer_batch_size = app.infer_batch_size
    c.blend, c.show_nrc, c.compact_inference = int(blend), int(show_nrc), int(compact_inference)
    c.x_begin, c.x_end = x_begin, x_end
    c.pipeline_train = int(pipeline_train)     # Train(N) underneath the tracking of frame N+1 (same order of effects)
    return c


def make_tile_render_config(width, height, app: AppConfig, rank: int, world: int, **kw) -> _lib.RenderConfig:
    """Render configuration of rank `rank` of `world` screen tiles: the frame's train-pixel lattice
    (CalcTrainSubset over the WHOLE frame) is cut into `world` column blocks; the rank owns the pixel columns that hold its block
    (the last rank also the remainder) and trains on its block only, i.e. on 1/world of every training batch -- `app` must be the
    per-rank AppConfig whose train batch size is the frame's divided by `world` (tile_app_config)."""
    t_frame = app.train_batch_count * app.train_batch_size * world
    c = make_render_config(width, height, app, train_pixels=t_frame, **kw)
    if c.train_width % world:
        raise ValueError(f"train lattice width {c.train_width} is not divisible by {world} tiles")
    tw = c.train_width // world
    c.train_tx0 = rank * tw
    c.x_begin = rank * tw * c.train_x_dist
    c.x_end = width if rank == world - 1 else (rank + 1) * tw * c.train_x_dist
    c.train_ring_size = int(app.train_ring_buf_size * float(tw * c.train_height))
    c.train_width = tw
    return c


def tile_app_config(app: AppConfig, world: int) -> AppConfig:
    """per-rank AppConfig of a tile-partitioned frame: 1/world of every training batch"""
    import copy
    shift = world.bit_length() - 1
    if world < 1 or (1 << shift) != world or app.log2_train_batch_size - shift < 7:
        raise ValueError("tile count must be a power of two that leaves training batches of >= 128 records")
    a = copy.deepcopy(app)
    a.log2_train_batch_size = app.log2_train_batch_size - shift
    return a


class NrcHpmRenderer:
    """``NrcHpmRenderer(width, height, blend, camera, appConfig, scene, nrc)``"""

    def __init__(self, width: int, height: int, blend: bool, camera: Camera, app_config: AppConfig, scene: HpmScene,
                 nrc: NeuralRadianceCache | None, *, render_config: _lib.RenderConfig | None = None, stream=None, **cfg_kw):
        self.width, self.height = width, height
        self.scene, self.nrc = scene, nrc
        self.cfg = render_config or make_render_config(width, height, app_config, blend=blend, **cfg_kw)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().hpm_renderer_create(scene._h, nrc._h if nrc is not None else None, C.byref(self.cfg), stream, C.byref(self._h)))
        self.SetCamera(camera)

    # ---- en::NrcHpmRenderer surface
    def Render(self, train: bool = True, frame_random=None):
        fr = np.asarray(frame_random if frame_random is not None else np.random.random(4), dtype=np.float32)
        _lib.check(_lib.lib().hpm_render(self._h, fr.ctypes.data_as(C.POINTER(C.c_float)), int(train)))

    def SetCamera(self, camera: Camera):
        self.camera = camera
        inv = np.ascontiguousarray(camera.inv_proj_view, np.float32)
        pos = np.asarray(camera.pos, np.float32)
        _lib.check(_lib.lib().hpm_renderer_set_camera(self._h, inv.ctypes.data_as(C.POINTER(C.c_float)), pos.ctypes.data_as(C.POINTER(C.c_float))))

    def SetBlend(self, blend: bool):
        _lib.check(_lib.lib().hpm_renderer_set_blend(self._h, int(blend)))

    def IsBlending(self) -> bool:
        return bool(self.cfg.blend)

    def EvaluateTimestampQueries(self) -> dict:
        ms = (C.c_float * 7)()
        _lib.check(_lib.lib().hpm_get_stage_ms(self._h, ms))
        return dict(zip(STAGE_NAMES, [float(v) for v in ms]))

    def GetFrameTimeMS(self) -> float:
        return self.EvaluateTimestampQueries()["total"]

    def GetImage(self) -> np.ndarray:
        """outputImage as float32 [H][W][4]"""
        return self.read(BUF_OUTPUT).reshape(self.height, self.width, 4)

    def ExportOutputImageToFile(self, file_path: str):
        """outputImage -> RGBA32F OpenEXR, the file tinyexr's SaveEXR writes (reference src/NrcHpmRenderer.cu:437-493)"""
        from . import exr
        exr.write_exr(file_path, self.GetImage())

    def Destroy(self):
        if self._h:
            _lib.check(_lib.lib().hpm_renderer_destroy(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass

    # ---- passes / buffers (parity tests, profiling)
    def _fr(self, frame_random):
        return np.asarray(frame_random, dtype=np.float32).ctypes.data_as(C.POINTER(C.c_float))

    def set_tracker_mode(self, mode: int):
        """0 automatic, 1 one pixel per thread, 2 path regeneration -- identical results, different schedules"""
        _lib.check(_lib.lib().hpm_renderer_set_tracker_mode(self._h, int(mode)))

    def pass_gen_rays(self, frame_random):
        fr = np.asarray(frame_random, dtype=np.float32)
        _lib.check(_lib.lib().hpm_pass_gen_rays(self._h, fr.ctypes.data_as(C.POINTER(C.c_float))))

    def pass_prep_train(self, frame_random):
        fr = np.asarray(frame_random, dtype=np.float32)
        _lib.check(_lib.lib().hpm_pass_prep_train(self._h, fr.ctypes.data_as(C.POINTER(C.c_float))))

    def pass_composite(self):
        _lib.check(_lib.lib().hpm_pass_composite(self._h))

    def mc_render(self, frame_random, path_length: int):
        fr = np.asarray(frame_random, dtype=np.float32)
        _lib.check(_lib.lib().hpm_mc_render(self._h, fr.ctypes.data_as(C.POINTER(C.c_float)), path_length))

    def sync(self):
        _lib.check(_lib.lib().hpm_sync(self._h))

    def buffer_info(self, which: int):
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(_lib.lib().hpm_buffer_info(self._h, which, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def read(self, which: int) -> np.ndarray:
        _, nbytes = self.buffer_info(which)
        dt = np.uint32 if which in (BUF_TRAIN_RING, BUF_INFER_FILTER) else np.uint64 if which == BUF_COUNTERS else np.float32
        out = np.empty(nbytes // np.dtype(dt).itemsize, dt)
        _lib.check(_lib.lib().hpm_read_buffer(self._h, which, out.ctypes.data_as(C.c_void_p), nbytes))
        return out

    def write(self, which: int, data: np.ndarray):
        data = np.ascontiguousarray(data)
        _lib.check(_lib.lib().hpm_write_buffer(self._h, which, data.ctypes.data_as(C.c_void_p), data.nbytes))


class McHpmRenderer(NrcHpmRenderer):
    """``McHpmRenderer(width, height, pathLength, blend, camera, scene)`` -- plain volumetric path tracer
    (reference src/McHpmRenderer.cpp, data/shader/mc/render.comp)."""

    def __init__(self, width, height, path_length, blend, camera, scene, *, stream=None):
        app = AppConfig.default()
        cfg = make_render_config(width, height, app, blend=blend, train_pixels=0)
        super().__init__(width, height, blend, camera, app, scene, None, render_config=cfg, stream=stream)
        self.path_length = path_length

    def Render(self, frame_random=None):
        fr = frame_random if frame_random is not None else np.random.random(4)
        self.mc_render(fr, self.path_length)
