"""Golden fixtures recorded from the reference's own code (tests/golden/make_golden.py) against the CPU oracle and against
the host-emulation build of the kernels; the GPU library is checked against the same files in tests/test_gpu_parity.py."""
import ctypes as C
import math
import os
import subprocess

import pytest

import golden_cases as GC
import orclib as O
import dumphfdl_b200.api as A

HERE = os.path.dirname(os.path.abspath(__file__))


class OraclePdu:
    """the oracle's record with the two levels in dB, as dispatch_pdu stores them (hfdl.c:1064-1065)"""

    def __init__(self, q):
        self.q = q
        self.rssi = 20.0 * math.log10(q.signal_level)
        self.noise_floor = 20.0 * math.log10(q.noise_floor)

    def __getattr__(self, n):
        return getattr(self.q, n)


@pytest.mark.parametrize("name", GC.FIXTURES)
def test_oracle_reproduces_the_reference_record(name):
    G = GC.load(name)
    raw = GC.capture_of(G)
    p = O.Pipeline(G["sample_rate"], GC.CF, G["freqs"], fold_mode=O.FOLD_SLICE, nthreads=4)
    p.feed(raw, O.SFMT_CS16)
    pdus = p.pdus()
    counters = {}
    for c, f in enumerate(G["freqs"]):
        a1, a2, m1, frames = p.stats(c)
        fr = [O.pdu_front(q.data()) for q in pdus if q.freq == f]
        counters[f] = {"demod.preamble.A2_found": a2, "demod.preamble.M1_found": m1, "demod.preamble.errors.M1_not_found": p.m1_not_found(c),
                       "frames.processed": frames, "frames.good": sum(v[0] == 0 for v in fr), "frame.errors.bad_fcs": sum(v[0] == 1 for v in fr),
                       "frame.errors.too_short": sum(v[0] == 2 for v in fr), "frame.dir.air2gnd": sum(v[0] == 0 and v[1] == 1 for v in fr),
                       "frame.dir.gnd2air": sum(v[0] == 0 and v[1] == 0 for v in fr), "lpdus.processed": sum(v[2] for v in fr),
                       "lpdus.good": sum(v[3] for v in fr), "lpdu.errors.bad_fcs": sum(v[4] for v in fr), "lpdu.errors.too_short": sum(v[5] for v in fr)}
    GC.check_against(G, [OraclePdu(q) for q in pdus], counters)


@pytest.fixture(scope="module")
def sim():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim")], check=True)
    return A.bind(C.CDLL(os.path.join(HERE, "cusim", "libhfdl_cusim.so")))


def test_host_emulation_of_the_kernels_reproduces_the_reference_record(sim, name="cfg1_multi.json"):
    # (the two-channel fixture's geometry and format are covered by test_cusim_logic.py::test_frontend_cfg1_cs16)
    G = GC.load(name)
    pdus, counters = GC.run_frontend(G, sim, batch=5)
    GC.check_against(G, pdus, counters)
