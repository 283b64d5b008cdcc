import sys, os, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from nrc_hpm_renderer_b200 import Camera, HpmSceneConfig, volume
from nrc_hpm_renderer_b200.renderer import HpmScene, McHpmRenderer
grid = volume.load_volume('/root/repo/data/wdas_cloud_quarter_u8.npz').data
refs = np.load('/root/repo/tests/golden/exr_block8.npz')
W, H = 240, 135
for sid, env in [(0,(0,0,0)),(1,(0,0,0)),(2,(0,0,0)),(4,(1,1,1)),(5,(1,1,1))]:
    ref = refs[f's{sid}'].astype(np.float32)
    for pl in (64, 1):
        scene = HpmScene(grid, HpmSceneConfig.preset(sid), env_color=env)
        r = McHpmRenderer(W, H, pl, True, Camera(aspect=1920/1080), scene)
        rng = np.random.default_rng(1337)
        for _ in range(128): r.Render(rng.random(4).astype(np.float32))
        img = r.GetImage()
        both = (ref[...,1] > 0.5) & (img[...,3] > 0.5)
        bg = (ref[...,1] == 0) & (img[...,3] == 0)
        print(f"scene {sid} pathlen {pl}: ours fg mean {img[...,0][both].mean():.5f} ref {ref[...,0][both].mean():.5f} ratio {img[...,0][both].mean()/ref[...,0][both].mean():.3f} alpha ours {img[...,3][both].mean():.4f} ref {ref[...,1][both].mean():.4f} bg ours {img[...,0][bg].mean():.4f} fgfrac {(img[...,3]>0.02).mean():.4f} vs {(ref[...,1]>0.02).mean():.4f}")
        r.Destroy(); scene.Destroy()
