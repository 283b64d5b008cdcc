"""Development tool: gen_rays pass at 1080p on the bundled cloud, pixel-per-thread kernel vs the path-regeneration schedule, for
BASELINE config 2 (scene 0, primaryRayLength 1) and config 4 (scene 5, primaryRayLength 4, primaryRayProb .75).
NRCHPM_WF_BLOCKS_PER_SM / NRCHPM_WF_ROUNDS / NRCHPM_WF_SPILL_BELOW are read by the library when the wavefront buffers are first allocated, so each setting is a fresh process:
    python scripts/tune_wavefront.py <mode> <config> [frames]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrc_hpm_renderer_b200 import AppConfig, Camera, HpmSceneConfig, volume
from nrc_hpm_renderer_b200 import renderer as R
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mode, config = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 20
W, H = 1920, 1080
grid = volume.load_volume(os.path.join(ROOT, "data", "wdas_cloud_quarter_u8.npz")).data
app = AppConfig.default()
if config == 4:
    app.scene = HpmSceneConfig.preset(5); app.primary_ray_length, app.primary_ray_prob = 4, 0.75
else:
    app.scene = HpmSceneConfig.preset(0)
scene = R.HpmScene(grid, app.scene)
r = R.NrcHpmRenderer(W, H, False, Camera(aspect=W / H), app, scene, None, render_config=R.make_render_config(W, H, app, train_pixels=0))
r.set_tracker_mode(mode)
rng = np.random.default_rng(1337)
frs = [rng.random(4).astype(np.float32) for _ in range(frames + 3)]
for fr in frs[:3]:
    r.pass_gen_rays(fr)
r.sync()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(frames + 1)]
ev[0].record()
for i, fr in enumerate(frs[3:]):
    r.pass_gen_rays(fr); ev[i + 1].record()
torch.cuda.synchronize()
ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(frames))
cnt = r.read(R.BUF_COUNTERS)
print(json.dumps({"mode": mode, "config": config, "blocks_per_sm": os.environ.get("NRCHPM_WF_BLOCKS_PER_SM"), "rounds": os.environ.get("NRCHPM_WF_ROUNDS"), "spill_below": os.environ.get("NRCHPM_WF_SPILL_BELOW"), "ms_median": round(ms[len(ms) // 2], 4), "ms_min": round(ms[0], 4),
                  "lookups": int(cnt[0]), "active": int(cnt[2])}))
