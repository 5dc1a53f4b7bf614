"""The front parser (which frames and LPDUs are "CRC-good", which statsd counters fire) pinned
against the REFERENCE'S OWN pdu.c / mpdu.c / spdu.c / lpdu.c / util.c / crc.c: those files are compiled where they lie
(oracle/Makefile -> oracle/_ref/libref_front.so; third-party headers are declaration-only stand-ins) and the reference's
real pdu_decoder_thread (pdu.c:91-178) parses every PDU; its statsd hooks and the protocol nodes it hands to a formatter
are the ground truth for orc_pdu_front_parse (and, through it, for the device's pdu_front_parse, tests/b200_cases.py).
Needs /root/reference (or a prebuilt oracle/_ref): skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

import orclib as O
from pdu_forge import Forge

pytestmark = pytest.mark.skipif(O.reflib_front() is None, reason="oracle/_ref/libref_front.so not built (reference tree absent)")


def fcs(buf):
    """FCS octets as the reference checks them (pdu.c:68-79): crc16_ccitt(buf, len, 0xFFFF) ^ 0xFFFF, little-endian -- computed
    by the reference's own crc.c"""
    a = np.frombuffer(bytes(buf), np.uint8).copy()
    v = O.reflib_front().crc16_ccitt(a, a.size, 0xFFFF) ^ 0xFFFF
    return bytes([v & 0xFF, v >> 8])


def expect_from_oracle(p):
    st, d, n, good, bad, short, mask = O.pdu_front(p)
    e = {"frames.processed": 1, "frames.good": int(st == 0), "frame.errors.bad_fcs": int(st == 1), "frame.errors.too_short": int(st == 2),
         "frame.dir.air2gnd": int(st == 0 and d == 1), "frame.dir.gnd2air": int(st == 0 and d == 0),
         "lpdus.processed": n, "lpdus.good": good, "lpdu.errors.bad_fcs": bad, "lpdu.errors.too_short": short, "other": 0}
    assert bin(mask).count("1") == min(good, 64) or n > 64
    return e


def check(pdus, **kw):
    ref = O.ref_front_run(pdus, **kw)
    for p, r in zip(pdus, ref):
        e = expect_from_oracle(p)
        assert {k: r[k] for k in e} == e, (len(p), bytes(p[:12]).hex(), r, e)
    return ref


def test_reference_fcs_check_and_transmitter_pdus():
    R = O.reflib_front()
    L = O.lib()
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
    L.orc_fcs_check.argtypes = [u8p, C.c_uint32]
    pd = [O.make_pdu(m, k, 7 + m) for m in range(8) for k in range(5)]
    hits = 0
    for p in pd:                                     # hfdl_pdu_fcs_check (pdu.c:68-79) == the oracle's, on good and on arbitrary octets
        a = np.frombuffer(p, np.uint8).copy()
        for n in (1, 6, 8, 9, 64, len(p) - 2):
            r = R.ref_front_fcs_check(a, n)
            assert r == L.orc_fcs_check(a, n)
            hits += r
    assert hits >= 16
    # "CRC-good" of a whole PDU as dispatch_pdu's consumers see it (orc_pdu_crc_good) == the reference's frames.good
    ref = check(pd + [b"\x03", b"\x00" * 10, b"\x13\x05", b"\x00" * 65, b"\x00" * 64 + fcs(b"\x00" * 64)])
    for p, r in zip(pd, ref):
        a = np.frombuffer(p, np.uint8).copy()
        assert bool(L.orc_pdu_crc_good(a, a.size)) == bool(r["frames.good"])
    # the transmitter's five kinds hit: good downlink, good SPDU, random octets, uplink with a bad LPDU, short / truncated LPDUs
    assert sum(r["frames.good"] for r in ref) >= 8 * 3 and any(r["lpdu.errors.bad_fcs"] for r in ref) and any(r["lpdu.errors.too_short"] for r in ref)
    # what reaches a formatter with the default configuration: good SPDUs, good LPDUs; never the MPDU node itself
    for p, r in zip(pd, ref):
        assert r["nodes.mpdu"] == 0 and r["nodes.other"] == 0
        assert r["nodes.lpdu"] == r["lpdus.good"]
        assert r["nodes.spdu"] == int((p[0] & 1) == 0 and r["frames.good"] == 1)


def test_constructed_mpdus_every_branch():
    rng = np.random.default_rng(20)
    F = Forge(rng, fcs)
    pd = F.branches()
    ref = check(pd)
    assert ref[6]["lpdus.good"] == 15 and ref[12]["lpdus.good"] == 8 and ref[14]["lpdus.good"] == 30
    assert ref[8]["lpdus.processed"] == 1 and ref[15]["lpdus.processed"] == 2
    assert ref[16]["frame.errors.too_short"] == 1 and ref[17]["frame.errors.too_short"] == 1
    assert [r["nodes.spdu"] for r in ref[18:]] == [1, 0, 0, 1]
    # LPDU types with their own length rules (lpdu.c:152-196): the frame check is the same for all of them
    types = []
    for t in (0x0D, 0x1D, 0x2F, 0x3F, 0x4F, 0x5F, 0x6F, 0x8F, 0x9F, 0xBF, 0xDF, 0xFF, 0x00):
        for n in (3, 4, 5, 6, 9, 12):
            hdr = bytearray(F.downlink([n])[:7])
            types.append(bytes(hdr) + fcs(hdr) + F.lpdu(n, True, typ=t))
    check(types)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_mutated_and_random_pdus(seed):
    pd = Forge(np.random.default_rng(500 + seed), fcs).fuzz(400)
    ref = check(pd)
    assert sum(r["frames.good"] for r in ref) > 100 and sum(r["frame.errors.bad_fcs"] for r in ref) > 30 and sum(r["frame.errors.too_short"] for r in ref) > 5
    assert sum(r["lpdu.errors.bad_fcs"] for r in ref) > 50 and sum(r["lpdu.errors.too_short"] for r in ref) > 10


def test_output_mpdus_and_corrupted_pdus_options_do_not_change_the_counters():
    """--output-mpdus / --output-corrupted-pdus (mpdu.c:39-45,121-128, lpdu.c:197-201, spdu.c:104-108) change what is delivered,
    never what is counted"""
    F = Forge(np.random.default_rng(9), fcs)
    pd = [F.downlink([20, 30], good=[True, False]), F.downlink([8], hdr_good=False), F.uplink([[9], [2]]),
          O.make_pdu(1, 1, 3), O.make_pdu(1, 2, 3)]
    base = check(pd)
    with_mpdus = check(pd, output_mpdus=True)
    corrupted = check(pd, output_mpdus=True, output_corrupted=True)
    assert [r["nodes.mpdu"] for r in with_mpdus] == [1, 0, 1, 0, int(pd[4][0] & 1 and with_mpdus[4]["frames.good"])]
    assert [r["nodes.lpdu"] for r in base] == [r["lpdus.good"] for r in base]
    assert [r["nodes.lpdu"] for r in corrupted][:3] == [r["lpdus.processed"] for r in base][:3]
    assert corrupted[1]["nodes.mpdu"] == 1
