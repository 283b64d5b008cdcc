"""GPU: the stand-alone tcgen05 / TMEM building-block probes (tests/cuda/*.cu, built by __graft_entry__.build())."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,token", [("tc05_probe", "TC05_PROBE_OK"), ("tc05_probe2", "TC05_PROBE2_OK")])
def test_probe(name, token):
    exe = os.path.join(ROOT, "tests", "cuda", name)
    if not os.path.exists(exe):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", exe, exe + ".cu"], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and token in res.stdout, res.stdout + res.stderr
