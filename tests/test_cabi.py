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


def test_interop_import_fails_loudly_on_bad_handles(built):
    """Vulkan interop helpers (reference src/NrcHpmRenderer.cu:644-690, 700-821): no Vulkan device exists here, so only the error paths
    can run -- a bad descriptor, a zero size, and (with or without a GPU) a descriptor that is not an exported Vulkan allocation must
    come back as error codes with a message, never as a crash or a mapped pointer."""
    from nrc_hpm_renderer_b200 import _lib
    l = _lib.lib()
    h, p = C.c_void_p(), C.c_void_p()
    assert l.nrchpm_import_external_buffer(-1, 1024, C.byref(h), C.byref(p)) == _lib.ERR_INVALID and not p.value
    assert b"file descriptor" in l.nrchpm_last_error()
    assert l.nrchpm_import_external_buffer(0, 0, C.byref(h), C.byref(p)) == _lib.ERR_INVALID
    r, w = os.pipe()                                    # a valid descriptor that is not an exported device allocation
    try:
        rc = l.nrchpm_import_external_buffer(r, 4096, C.byref(h), C.byref(p))
        assert rc != 0 and not p.value and not h.value
        s = C.c_void_p()
        assert l.nrchpm_import_external_semaphore(-1, C.byref(s)) == _lib.ERR_INVALID
        rc = l.nrchpm_import_external_semaphore(w, C.byref(s))
        assert rc != 0 and not s.value
    finally:
        for fd in (r, w):
            try:
                os.close(fd)
            except OSError:
                pass                                    # (a failed import may have consumed it)
    assert l.nrchpm_release_external_buffer(None) == 0 and l.nrchpm_release_external_semaphore(None) == 0


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
