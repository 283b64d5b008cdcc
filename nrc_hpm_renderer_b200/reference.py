"""``en::Reference`` -- frame metrics against a converged reference frame (reference include/engine/graphics/Reference.hpp,
src/Reference.cpp; compute shaders data/shader/ref/cmp1.comp, norm.comp, cmp2.comp).

The reference image is `reference/<scene id>/0.exr` (RGBA32F OpenEXR): loaded if the folder exists, otherwise generated the way
the reference's (compiled-out) code path would -- `McHpmRenderer(width, height, 64, blend=True)` rendered 8192 times
from the fixed reference camera -- and exported.  `CompareNrc` / `CompareMc` move the renderer to the reference camera, render one
frame without training and reduce the two images on the GPU (`hpm_compare_images`); the numbers are the ones the reference logs
(`MSE | rBias | rVar`) and writes to its CSV (`mse, relBias, CV`).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass

import numpy as np

from . import _lib, exr
from .camera import Camera

REF_PATH_LENGTH, REF_FRAMES = 64, 8192          # src/Reference.cpp:581, 591


@dataclass
class Result:
    """Reference::Result (Reference.hpp:17-29)"""
    mse: float
    refMean: float
    ownMean: float
    ownVar: float
    validPixelCount: int

    def GetBias(self) -> float: return self.ownMean - self.refMean
    def GetRelBias(self) -> float: return self.GetBias() / self.refMean
    def GetRelVar(self) -> float: return self.ownVar / self.refMean
    def GetCV(self) -> float: return math.sqrt(self.ownVar) / self.ownMean


def compare_device_images(d_ref: int, d_cmp: int, width: int, height: int, stream=None) -> Result:
    """both arguments: device pointers to float[W*H][4] images"""
    r = _lib.CompareResult()
    _lib.check(_lib.lib().hpm_compare_images(d_ref, d_cmp, width, height, C.byref(r), stream))
    return Result(r.mse, r.ref_mean, r.own_mean, r.own_var, int(r.valid_pixel_count))


def ref_camera(width: int, height: int) -> Camera:
    """Reference::CreateRefCameras (src/Reference.cpp:443-455) -- Camera's defaults are exactly this pose"""
    return Camera(pos=(64.0, 0.0, 0.0), view_dir=(-1.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), aspect=width / height, fov=math.radians(60.0), near=0.1, far=100.0)


class Reference:
    """``Reference(width, height, appConfig, scene)``; `reference_root` replaces the working-directory-relative "reference/"."""

    def __init__(self, width: int, height: int, app_config, scene, *, reference_root: str = "reference", frames: int = REF_FRAMES,
                 path_length: int = REF_PATH_LENGTH, seed: int = 1337):
        import torch
        self.width, self.height = width, height
        self.m_RefCamera = ref_camera(width, height)
        ref_dir = os.path.join(reference_root, str(app_config.scene.id))
        path = os.path.join(ref_dir, "0.exr")
        if not os.path.isdir(ref_dir):                                        # Reference::GenRefImages (:568-606)
            from .renderer import McHpmRenderer
            os.makedirs(ref_dir)
            mc = McHpmRenderer(width, height, path_length, True, self.m_RefCamera, scene)
            rng = np.random.default_rng(seed)
            for _ in range(frames):
                mc.Render(rng.random(4).astype(np.float32))
            mc.ExportOutputImageToFile(path)
            mc.Destroy()
        img = exr.read_exr(path)                                              # LoadEXR (:617-627)
        if img.shape[0] != height or img.shape[1] != width:
            raise _lib.NrcHpmError(_lib.ERR_INVALID, f"{path} has wrong resolution")
        self.m_RefImage = torch.from_numpy(np.ascontiguousarray(img)).cuda()

    def _compare(self, renderer) -> Result:
        from .renderer import BUF_OUTPUT
        d_cmp, _ = renderer.buffer_info(BUF_OUTPUT)
        renderer.sync()
        return compare_device_images(self.m_RefImage.data_ptr(), d_cmp, self.width, self.height)

    def CompareNrc(self, renderer, old_camera, frame_random=None) -> Result:
        """Reference::CompareNrc (:72-112): reference camera, Render(train=false), compare, restore the camera"""
        renderer.SetCamera(self.m_RefCamera)
        renderer.Render(False, frame_random)
        res = self._compare(renderer)
        renderer.SetCamera(old_camera)
        return res

    def CompareMc(self, renderer, old_camera, frame_random=None) -> Result:
        """Reference::CompareMc (:114-155)"""
        renderer.SetCamera(self.m_RefCamera)
        renderer.Render(frame_random)
        res = self._compare(renderer)
        renderer.SetCamera(old_camera)
        return res

    def Destroy(self):
        self.m_RefImage = None
