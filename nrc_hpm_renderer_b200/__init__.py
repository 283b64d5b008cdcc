"""nrc_hpm_renderer_b200 -- B200-native (sm_100a) implementation of the NRC-HPM-Renderer hot path:
volumetric delta/ratio tracking with NRC path termination + Neural Radiance Cache inference and online training.

Host-side mirror of the reference interface (same names and argument meaning):
    AppConfig, HpmSceneConfig        reference include/engine/AppConfig.hpp, src/AppConfig.cpp
    Camera                           reference src/Camera.cpp (matrix math only)
    NeuralRadianceCache              reference include/engine/graphics/NeuralRadianceCache.hpp
    HpmScene, NrcHpmRenderer, McHpmRenderer
    Reference, Result                reference include/engine/graphics/Reference.hpp, src/Reference.cpp (frame metrics)
    exr.read_exr / exr.write_exr     the reference's on-disk frame format (tinyexr SaveEXR / LoadEXR)
The compute lives in libnrchpm_b200.so (C ABI: include/nrc_hpm_b200.h); there is no CPU fallback.
"""
from .config import AppConfig, HpmSceneConfig, calc_train_subset, encoding_json, sky_size  # noqa: F401
from .camera import Camera  # noqa: F401


def __getattr__(name):
    # the CUDA-backed classes import lazily so that config / camera / volume helpers work without the shared library
    if name in ("NeuralRadianceCache",):
        from . import nrc
        return getattr(nrc, name)
    if name in ("HpmScene", "NrcHpmRenderer", "McHpmRenderer", "make_render_config"):
        from . import renderer
        return getattr(renderer, name)
    if name in ("Reference", "Result"):
        from . import reference
        return getattr(reference, name)
    raise AttributeError(name)
