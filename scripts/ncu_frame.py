"""Development tool: a few full frames on the bundled cloud, for ncu captures of the tracking kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig, volume
from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
from nrc_hpm_renderer_b200.renderer import HpmScene, NrcHpmRenderer
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 1920, 1080
grid = volume.load_volume(os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")).data
app = AppConfig.default(); app.scene = HpmSceneConfig.preset(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
nrc = NeuralRadianceCache(app); scene = HpmScene(grid, app.scene)
r = NrcHpmRenderer(W, H, False, Camera(aspect=W / H), app, scene, nrc)
rng = np.random.default_rng(1337)
for i in range(6):
    r.Render(True, rng.random(4).astype(np.float32))
r.sync()
print(r.EvaluateTimestampQueries())
