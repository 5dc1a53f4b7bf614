"""The C-ABI library must load without a GPU and export every symbol include/*.h declares; it must refuse to
create a front-end when no CUDA device exists (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dumphfdl_b200", "libhfdl_b200.so")


def declared_symbols():
    names = set()
    for h in ("hfdl_b200.h", "hfdl_b200_block.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b((?:hfdl_b200|hfdl_gpu|cbuffercf)_\w+)\s*\(", txt))
    names -= {"hfdl_gpu_pdu_callback"}
    return sorted(names)


@pytest.mark.skipif(not os.path.exists(LIB), reason="libhfdl_b200.so not built: run __graft_entry__.build()")
def test_library_loads_and_exports_every_declared_symbol():
    L = C.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


@pytest.mark.skipif(not os.path.exists(LIB), reason="libhfdl_b200.so not built")
def test_no_cpu_fallback_without_a_device():
    import dumphfdl_b200 as hb
    L = hb.load()
    if L.hfdl_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        hb.Frontend(250000, 10000000, [10063000])
    import numpy as np
    with pytest.raises(RuntimeError):
        hb.fft_forward(np.zeros(64, np.complex64))


@pytest.mark.skipif(not os.path.exists(LIB), reason="libhfdl_b200.so not built")
def test_library_does_not_define_host_program_symbols():
    """The ring (liquid-dsp's cbuffercf) and the PDU hand-off belong to the host program: the library only holds weak
    references to them and must not define (and so interpose) any of them."""
    import subprocess
    out = subprocess.run(["nm", "-D", LIB], capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in out.splitlines()]
    for name in ("cbuffercf_size", "cbuffercf_read", "cbuffercf_release", "pdu_decoder_queue_push", "hfdl_pdu_metadata_create", "octet_string_new"):
        kinds = [r[-2] for r in rows if r and r[-1] == name]
        assert kinds and all(k in ("w", "U") for k in kinds), (name, kinds)
    defined = [r[-1] for r in rows if len(r) == 3 and r[1] in ("T", "D", "B") and r[-1].startswith("cbuffercf")]
    assert not defined, defined
    # and nothing of the test-side oracle is linked in
    assert not any("orc_" in r[-1] for r in rows if r)
