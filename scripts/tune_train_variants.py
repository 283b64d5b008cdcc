"""Development tool: training-step time (2^14 records, hash grid, 64 x 6) of the product library and of every variant under
nrc_hpm_renderer_b200/variants/ (each in its own process: NRCHPM_LIB)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = "import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import tune_train; r = tune_train.run(1, 2, sizes=(1 << 14, 1 << 20)); print(json.dumps(r))" % (ROOT, os.path.join(ROOT, "scripts"))
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "nrc_hpm_renderer_b200", "variants", "*.so"))):
    env = dict(os.environ)
    if lib: env["NRCHPM_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    print((lib or "product").split("/")[-1], r.stdout.strip()[:330] or r.stderr.strip()[-300:], flush=True)
