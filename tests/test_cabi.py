"""CPU: the C-ABI shared library loads and exports every symbol include/nrc_hpm_b200.h declares; without a GPU the
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "nrc_hpm_b200.h")
LIB = os.path.join(ROOT, "nrc_hpm_renderer_b200", "libnrchpm_b200.so")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:nrc|hpm|nrchpm)_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "nrc_hpm_renderer_b200", "csrc")], check=True, capture_output=True)
    return LIB


def test_every_declared_symbol_is_exported(built):
    names = declared_functions()
    assert len(names) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [n for n in names if n not in exported]
    assert not missing, missing


def test_python_binding_covers_the_header(built):
    from nrc_hpm_renderer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    l = _lib.lib()
    assert l.nrchpm_version() >= 100


def test_sm100a_code_is_in_the_library(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True)
    assert "sm_100a" in out.stdout


def test_fails_loudly_without_gpu_or_with_bad_config(built):
    import torch
    from nrc_hpm_renderer_b200 import _lib
    l = _lib.lib()
    h = C.c_void_p()
    rc = l.nrc_create(b'{"encoding": {"otype": "Composite", "nested": [{"otype": "Nope"}, {"otype": "OneBlob"}]}}', 1337, C.byref(h))
    assert rc == _lib.ERR_UNSUPPORTED and b"not one of the reference presets" in l.nrchpm_last_error()
    rc = l.nrc_create(b"{not json", 1337, C.byref(h))
    assert rc == _lib.ERR_INVALID
    if not torch.cuda.is_available():
        cfg = b'{"encoding": {"otype": "Composite", "nested": [{"otype": "Identity"}, {"otype": "Identity"}]}, "network": {"n_hidden_layers": 2}}'
        rc = l.nrc_create(cfg, 1337, C.byref(h))
        assert rc == _lib.ERR_CUDA and not h.value                     # no device -> error, never a CPU path
        with pytest.raises(_lib.NrcHpmError):
            from nrc_hpm_renderer_b200.nrc import NeuralRadianceCache
            NeuralRadianceCache()
