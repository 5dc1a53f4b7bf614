"""Pins the hfdl.c half of the oracle against the REFERENCE'S OWN hfdl.c.

oracle/_ref/libref.so holds /root/reference/src/{block,fft,fastddc,libcsdr,libcsdr_gpl,hfdl,input-helpers,crc}.c and
libfec/viterbi27_port.c compiled where they lie (oracle/Makefile), wired as main.c:697-774 wires them and fed as
input-file.c:50-63 feeds them (oracle/ref_shim/ref_host.c).  liquid-dsp is not installed anywhere reachable, so the
liquid symbols are served by the oracle's own restated objects (oracle/orc_liquid.c through ref_shim/liquid_shim.c):
the reference's sample loop, Costas loop, sampler, framer FSM, descrambler, deinterleaver, decode_user_data and
dispatch_pdu (hfdl.c:250-413,593-1080) then run UNMODIFIED on exactly the object arithmetic the oracle uses, and
every DATADUMPS tap (hfdl.c:616-655) and every PDU must be IDENTICAL to the oracle's restatement of hfdl.c
(oracle/orc_hfdl.c).  What stays unpinned is the inside of the liquid objects themselves (stated in oracle/orc.h).

Skipped when oracle/_ref/libref.so is absent and /root/reference is not there to build it from (the GPU box has the
prebuilt file; nothing here reads /root/reference at run time)."""
import numpy as np
import pytest

import b200_cases as K
import orclib as O

pytestmark = pytest.mark.skipif(O.reflib() is None or not hasattr(O.reflib(), "ref_pipeline_create"),
                                reason="oracle/_ref/libref.so (reference sources compiled in place) not available")

SR, CF, F = 250000, 10000000, 10063000
TAPS = [("f_chan_out.cf32", "chan"), ("f_agc_out.cf32", "agc"), ("f_mf_out.cf32", "mf"), ("f_symsync_out.cf32", "symsync"),
        ("f_costas_out.cf32", "costas"), ("f_eq_out.cf32", "eq")]
BITRATE = {0: 300, 1: 600, 2: 1200, 3: 1800}


def run_both(sr, freqs, raw, sfmt=O.SFMT_CF32, taps=True, fft_threads=2):
    r = O.RefPipeline(sr, CF, freqs, sfmt=sfmt, fft_threads=fft_threads, dumps=taps)
    r.feed(raw)
    r.finish()
    p = O.Pipeline(sr, CF, freqs, fold_mode=O.FOLD_FULL, nthreads=2)
    if taps:
        for c in range(len(freqs)):
            p.set_capture(c, [t for _, t in TAPS], 1 << 21)
    p.feed(raw, sfmt)
    return r, p


def assert_same_pdus(r, p, freqs):
    rp, op = r.pdus(), p.pdus()
    assert len(rp) == len(op)
    for f in freqs:      # the reference delivers in thread-race order across channels; per channel the order is fixed
        a = [q for q in rp if q.freq == f]
        b = [q for q in op if q.freq == f]
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x.data() == y.data()                                     # octets identical, CRC-failed ones included
            assert x.version == 1 and x.flags == 0                          # hfdl.c:1061,1075
            assert (x.bit_rate, x.slot) == (y.bit_rate, y.slot) == (BITRATE[y.M1 % 4], b"S" if y.M1 < 4 else b"D")
            # same float expressions on the same inputs (hfdl.c:812,1063-1065): identical, not merely close
            assert x.freq_err_hz == y.freq_err_hz
            assert x.rssi == np.float32(20.0) * np.log10(np.float32(y.signal_level), dtype=np.float32) or abs(x.rssi - 20 * np.log10(y.signal_level)) < 1e-4
            assert abs(x.noise_floor - 20 * np.log10(y.noise_floor)) < 1e-4
    return rp, op


def assert_same_taps(r, p):
    # the reference opens one dump per channel thread under the same name, in thread-start order: every oracle
    # channel must find its bit-identical twin among them
    nch = len(p.freqs)
    for name, tap in TAPS:
        dumps = [r.dump(name, i)[1] for i in range(nch)]
        assert all(v is not None and v.size > 0 for v in dumps), name
        left = list(range(nch))
        for c in range(nch):
            o = p.capture(c, tap).view(np.uint32)
            hit = [i for i in left if dumps[i].size * 2 == o.size and np.array_equal(dumps[i].view(np.uint32), o)]
            assert hit, (name, c)                                              # bit-identical floats
            left.remove(hit[0])


def one_frame(M1, seed, esn0=20.0, amp=0.1, cfo=7.0):
    pdu = O.make_pdu(M1, kind=M1 % 2, seed=seed)
    dur = 5.6 if M1 >= 4 else 3.2
    fr = O.tx_frame(F, M1, 0.25, pdu, cfo_hz=cfo, phase0=1.0, amplitude=amp)
    return pdu, O.render(int(SR * dur), SR, CF, [fr], noise_sigma=O.noise_sigma(amp, SR, esn0), seed=seed)


@pytest.mark.parametrize("M1", range(8))
def test_reference_hfdl_equals_oracle_all_modes(M1):
    pdu, x = one_frame(M1, 50 + M1)
    r, p = run_both(SR, [F], x)
    rp, op = assert_same_pdus(r, p, [F])
    assert len(rp) == 1 and rp[0].data() == pdu                               # and it is what was transmitted
    assert_same_taps(r, p)
    a1, a2, m1, frames = p.stats(0)
    assert r.stat(F, "demod.preamble.A2_found") == a2 and r.stat(F, "demod.preamble.M1_found") == m1     # hfdl.c:818,828
    assert r.stat(F, "demod.preamble.errors.M1_not_found") == p.m1_not_found(0)                          # hfdl.c:840
    r.close()
    p.close()


@pytest.mark.parametrize("esn0,M1", [(8.0, 3), (5.0, 1), (5.0, 2), (3.0, 0)])
def test_reference_hfdl_equals_oracle_low_snr(esn0, M1):
    # bit errors, failed preamble searches and framer resets take the rarely used branches of hfdl.c:779-891
    pdu, x = one_frame(M1, 70 + M1, esn0=esn0)
    r, p = run_both(SR, [F], x)
    assert_same_pdus(r, p, [F])
    assert_same_taps(r, p)
    assert r.stat(F, "demod.preamble.errors.M1_not_found") == p.m1_not_found(0)
    r.close()
    p.close()


def test_reference_hfdl_noise_only_resets_and_counters():
    # noise with a strong off-tune carrier: A1 false alarms, A2 retries, M1_not_found, Costas blow-ups (hfdl.c:711-715)
    rng = np.random.default_rng(5)
    n = SR * 6
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.02).astype(np.complex64)
    t = np.arange(n) / SR
    x += (0.05 * np.exp(2j * np.pi * (F + 1440 - CF + 400.0) * t)).astype(np.complex64)
    r, p = run_both(SR, [F], x)
    assert_same_pdus(r, p, [F])
    assert_same_taps(r, p)
    a1, a2, m1, frames = p.stats(0)
    assert r.stat(F, "demod.preamble.A2_found") == a2 and r.stat(F, "demod.preamble.M1_found") == m1
    assert r.stat(F, "demod.preamble.errors.M1_not_found") == p.m1_not_found(0)
    r.close()
    p.close()


@pytest.mark.parametrize("name", K.HOSTILE)
def test_reference_hfdl_equals_oracle_on_hostile_captures(name):
    """collisions, carrier offsets beyond the loop's comfort, clipping, frames cut by the capture's edges, a strong adjacent
    carrier, a DC spur: whatever the reference's demodulator makes of them, the oracle makes the same -- every tap
    bit-identical, the same PDUs (none, or with bit errors behind a good or bad FCS), the same counters"""
    x = K.hostile_capture(name, SR, F)
    r, p = run_both(SR, [F], x)
    rp, _ = assert_same_pdus(r, p, [F])
    assert_same_taps(r, p)
    a1, a2, m1, frames = p.stats(0)
    assert r.stat(F, "demod.preamble.A2_found") == a2 and r.stat(F, "demod.preamble.M1_found") == m1
    assert r.stat(F, "demod.preamble.errors.M1_not_found") == p.m1_not_found(0)
    assert len(rp) == frames
    r.close()
    p.close()


@pytest.mark.parametrize("seed", range(16))
def test_reference_hfdl_equals_oracle_on_random_jobs(seed):
    """differential test: seeded random jobs (channels, frames per channel, modes, PDU kinds, carrier offsets, amplitudes,
    slot gaps / overlaps, Es/N0 3-25 dB) through the reference's own code and through the oracle"""
    freqs, x = K.random_scenario(seed, SR)
    r, p = run_both(SR, freqs, x)
    rp, _ = assert_same_pdus(r, p, freqs)
    assert_same_taps(r, p)
    for c, f in enumerate(freqs):
        a1, a2, m1, frames = p.stats(c)
        assert r.stat(f, "demod.preamble.A2_found") == a2 and r.stat(f, "demod.preamble.M1_found") == m1
        assert r.stat(f, "demod.preamble.errors.M1_not_found") == p.m1_not_found(c)
    r.close()
    p.close()


def test_reference_multichannel_back_to_back_and_raw_formats():
    sr = 250000
    freqs = [10021000, 10063000, 10090000]
    pd = [O.make_pdu(1, 0, 41), O.make_pdu(3, 1, 42), O.make_pdu(0, 0, 43), O.make_pdu(6, 1, 44)]
    frames = [O.tx_frame(freqs[1], 1, 0.2, pd[0], cfo_hz=-12, amplitude=0.05),
              O.tx_frame(freqs[1], 3, 0.2 + 2.4615, pd[1], cfo_hz=9, phase0=2, amplitude=0.05),
              O.tx_frame(freqs[0], 0, 0.9, pd[2], cfo_hz=3, amplitude=0.08),
              O.tx_frame(freqs[2], 6, 0.4, pd[3], cfo_hz=-5, amplitude=0.06)]
    x = O.render(int(sr * 6.0), sr, CF, frames, noise_sigma=O.noise_sigma(0.05, sr, 15.0), seed=9)
    r, p = run_both(sr, freqs, x, taps=False)
    rp, _ = assert_same_pdus(r, p, freqs)
    assert {(q.freq, q.data()) for q in rp} == {(freqs[1], pd[0]), (freqs[1], pd[1]), (freqs[0], pd[2]), (freqs[2], pd[3])}
    r.close()
    p.close()
    # CS16 / CU8 through the reference's own convert_cs16 / convert_cu8 (input-helpers.c:37-78) vs orc_convert_samples
    n = x.size
    s16 = np.zeros(2 * n, np.int16)
    O.lib().orc_quantize_cs16(x, n, s16)
    r, p = run_both(sr, freqs, s16, sfmt=O.SFMT_CS16, taps=True)
    assert_same_pdus(r, p, freqs)
    assert_same_taps(r, p)
    r.close()
    p.close()
    u8 = np.zeros(2 * n, np.uint8)
    O.lib().orc_quantize_cu8((x * 2).astype(np.complex64), n, u8)
    r, p = run_both(sr, freqs, u8, sfmt=O.SFMT_CU8, taps=True)
    assert_same_pdus(r, p, freqs)
    assert_same_taps(r, p)
    r.close()
    p.close()


def test_reference_cfg2_geometry_two_megasamples():
    # BASELINE config 2 geometry (2 Msps, N = 262144): reference fft.c + block.c + fastddc.c + hfdl.c end to end
    sr = 2000000
    freqs = [9700000, 10063000, 10400000]
    pd = [O.make_pdu(m, m % 2, 90 + m) for m in (1, 2, 3)]
    frames = [O.tx_frame(f, m, 0.1 + 0.05 * i, pd[i], cfo_hz=4.0 * i - 3, phase0=i, amplitude=0.08) for i, (f, m) in enumerate(zip(freqs, (1, 2, 3)))]
    x = O.render(int(sr * 2.9), sr, CF, frames, noise_sigma=O.noise_sigma(0.08, sr, 18.0), seed=12)
    r, p = run_both(sr, freqs, x, taps=True)
    rp, _ = assert_same_pdus(r, p, freqs)
    assert sorted(q.data() for q in rp) == sorted(pd)
    assert_same_taps(r, p)
    r.close()
    p.close()


@pytest.mark.parametrize("sr", [768000, 1024000, 2400000, 5000000])
def test_reference_other_sample_rates(sr):
    """geometries between the BASELINE configurations (fastddc_init picks M = 2048 or 4096 and resampler rates from 0.55 to
    0.9 depending on the rate): reference end to end == oracle, taps bit-identical"""
    freqs = [CF - int(0.41 * sr) // 1000 * 1000, CF + 63000]
    pd = [O.make_pdu(m, m % 2, 190 + m) for m in (2, 5)]
    frames = [O.tx_frame(f, m, 0.1 + 0.05 * i, pd[i], cfo_hz=5.0 * i - 3, phase0=i, amplitude=0.08) for i, (f, m) in enumerate(zip(freqs, (2, 5)))]
    x = O.render(int(sr * 5.4), sr, CF, frames, noise_sigma=O.noise_sigma(0.08, sr, 18.0), seed=14)
    r, p = run_both(sr, freqs, x, taps=True)
    rp, _ = assert_same_pdus(r, p, freqs)
    assert sorted(q.data() for q in rp) == sorted(pd)
    assert_same_taps(r, p)
    r.close()
    p.close()


def test_reference_cfg3_geometry_twenty_megasamples_cs16():
    # BASELINE config 3 geometry as bench.py runs it (20 Msps CS16: N = 2^22, M = 4096, 1792 outputs per block, resampler
    # 0.55296): the reference's input-helpers.c conversion + fft.c + fastddc.c + hfdl.c end to end vs the oracle
    sr = 20000000
    freqs = [CF - 8123000, CF + 4660000]
    pd = [O.make_pdu(m, m % 2, 190 + m) for m in (1, 3)]
    frames = [O.tx_frame(f, m, 0.05 + 0.01 * i, pd[i], cfo_hz=6.0 * i - 3, phase0=0.7 * i, amplitude=0.1) for i, (f, m) in enumerate(zip(freqs, (1, 3)))]
    x = O.render(int(sr * 2.8), sr, CF, frames, noise_sigma=O.noise_sigma(0.1, sr, 20.0), seed=14)
    raw = np.zeros(2 * x.size, np.int16)
    O.lib().orc_quantize_cs16(x, x.size, raw)
    del x
    r, p = run_both(sr, freqs, raw, sfmt=O.SFMT_CS16, taps=True)
    assert (p.ddc.fft_size, p.ddc.fft_inv_size, p.ddc.input_size, p.ddc.post_input_size // p.ddc.post_decimation) == (1 << 22, 4096, 3670016, 1792)
    rp, _ = assert_same_pdus(r, p, freqs)
    assert sorted(q.data() for q in rp) == sorted(pd)
    assert_same_taps(r, p)
    r.close()
    p.close()


@pytest.mark.parametrize("sr,N,sfmt", [(30000000, 1 << 22, O.SFMT_CF32), (60000000, 1 << 23, O.SFMT_CS16)])
def test_reference_cfg4_cfg5_geometry(sr, N, sfmt):
    # BASELINE config 4 / 5 geometries (30 Msps: N = 2^22, M = 2048; 60 Msps: N = 2^23, overlap 2^20): reference vs oracle
    freqs = [CF - 9100000, CF + 3907000]
    pd = [O.make_pdu(m, m % 2, 290 + m) for m in (2, 0)]
    frames = [O.tx_frame(f, m, 0.04 + 0.01 * i, pd[i], cfo_hz=5.0 * i - 2, phase0=0.3 * i, amplitude=0.1) for i, (f, m) in enumerate(zip(freqs, (2, 0)))]
    x = O.render(int(sr * 2.7), sr, CF, frames, noise_sigma=O.noise_sigma(0.1, sr, 20.0), seed=15)
    raw = x
    if sfmt == O.SFMT_CS16:
        raw = np.zeros(2 * x.size, np.int16)
        O.lib().orc_quantize_cs16(x, x.size, raw)
        del x
    r, p = run_both(sr, freqs, raw, sfmt=sfmt, taps=True)
    assert (p.ddc.fft_size, p.ddc.fft_inv_size) == (N, 2048)
    rp, _ = assert_same_pdus(r, p, freqs)
    assert sorted(q.data() for q in rp) == sorted(pd)
    assert_same_taps(r, p)
    r.close()
    p.close()


def test_scrambler_same_under_both_liquid_msequence_conventions():
    # hfdl.c:333-345 picks (genpoly, init) by liquid version so that both conventions emit one sequence; the oracle's
    # orc_scrambler_bits must be that sequence (first 32 bits recorded in SURVEY appendix A)
    import ctypes as C
    L = O.lib()

    class MS(C.Structure):
        _fields_ = [("m", C.c_uint32), ("g", C.c_uint32), ("a", C.c_uint32), ("v", C.c_uint32), ("convention", C.c_int)]
    L.orc_msequence_init.argtypes = [C.POINTER(MS), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    L.orc_msequence_advance.argtypes = [C.POINTER(MS)]
    L.orc_msequence_advance.restype = C.c_uint32
    old, new = MS(), MS()
    L.orc_msequence_init(C.byref(old), 15, 0x8002, 0x6959, 0)       # liquid < 1.6.0
    L.orc_msequence_init(C.byref(new), 15, 0x4001, 0x4d4b, 1)       # liquid >= 1.6.0
    a = [L.orc_msequence_advance(C.byref(old)) for _ in range(120)]
    b = [L.orc_msequence_advance(C.byref(new)) for _ in range(120)]
    scr = np.zeros(240, np.uint8)
    L.orc_scrambler_bits(scr, 240)
    assert a == b == list(scr[:120]) == list(scr[120:])
    assert "".join(map(str, a[:32])) == "01100011001000110111101110000100"


def test_block_struct_layout_matches_reference_headers():
    # include/hfdl_b200_block.h restates the ABI of src/block.h:27-68 and src/pdu.h:8-17: sizes and offsets as the
    # reference's own headers compile here
    import subprocess, tempfile, os
    R = O.reflib()
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "hfdl_b200_block.h"
int main(void) { printf("%zu %zu %zu %zu\n", sizeof(struct block), sizeof(struct block_connection), offsetof(struct block, thread_routine), offsetof(struct block, running)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(O.ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()
    assert [int(v) for v in out] == [R.ref_struct_sizes(0), R.ref_struct_sizes(1), R.ref_struct_sizes(3), R.ref_struct_sizes(5)]
