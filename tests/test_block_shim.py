"""The block.c-facing wrapper (include/hfdl_b200_block.h): tests/block_driver.c wires an input block, the one2one
ring and the GPU front-end block exactly as dumphfdl's main.c / block.c do, replays a capture file through it and
prints the PDUs.  CPU variant uses the host-emulation build; the GPU variant the real library."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import b200_cases as K
import orclib as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def run_driver(libpath, sr, freqs, modes, dur, seed):
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim"), "all"], check=True)
    x, truth = K.make_capture(sr, freqs, modes, dur, seed=seed)
    ref = K.run_oracle(sr, freqs, x, O.SFMT_CF32).pdus()
    with tempfile.NamedTemporaryFile(suffix=".cf32", delete=False) as f:
        x.tofile(f)
        path = f.name
    try:
        out = subprocess.run([os.path.join(HERE, "block_driver"), libpath, path, str(sr), str(K.CF)] + [str(f) for f in freqs],
                             capture_output=True, text=True, timeout=900, check=True).stdout
    finally:
        os.unlink(path)
    got = sorted((int(a[1]), int(a[2]), int(a[3]), int(a[4]), a[5]) for a in (l.split() for l in out.splitlines()) if a and a[0] == "PDU")
    want = sorted((q.freq, q.M1, q.crc_good, int(q.sample_cnt_a2), q.data().hex()) for q in ref)
    assert got == want
    assert sorted((g[0], bytes.fromhex(g[4])) for g in got) == sorted(truth)
    return len(got)


def test_block_contract_host_emulation():
    lib = os.path.join(HERE, "cusim", "libhfdl_cusim.so")
    assert run_driver(lib, 250000, [10063000, 9952000], [1, 2], 3.3, seed=31) == 2


@pytest.mark.gpu
def test_block_contract_gpu():
    lib = os.path.join(ROOT, "dumphfdl_b200", "libhfdl_b200.so")
    assert run_driver(lib, 2000000, [K.CF + 212000, K.CF - 424000, K.CF + 636000], [3, 5, 0], 5.8, seed=33) == 3
