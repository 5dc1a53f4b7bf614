"""Pins the oracle (oracle/*.c) against (1) the reference's own sources compiled in place into
oracle/_ref/libref.so -- fastddc.c, libcsdr.c, libcsdr_gpl.c, crc.c, libfec/viterbi27_port.c -- and
(2) the known answers the reference text holds.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import orclib as O

L = O.lib()
R = O.reflib()
needs_ref = pytest.mark.skipif(R is None, reason="oracle/_ref/libref.so not built (reference tree absent)")

RATES = [250000, 2000000, 20000000, 30000000, 60000000, 12000, 768000, 10000, 2400000]


def test_fft_matches_numpy():
    rng = np.random.default_rng(1)
    for n in [2, 4, 8, 32, 512, 2048, 4096, 32768, 262144]:
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        for d in (1, -1):
            y = np.zeros(n, np.complex64)
            L.orc_fft(x, y, n, d)
            ref = np.fft.fft(x.astype(np.complex128)) if d == 1 else np.fft.ifft(x.astype(np.complex128)) * n
            assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 5e-7


def test_geometry_table_survey8():
    # SURVEY.md section 8 table (fastddc_init evaluated in float32)
    want = {250000: (32, 32768, 4096, 28672, 2048, 256, 1792), 2000000: (256, 262144, 32768, 229376, 2048, 256, 1792),
            20000000: (2048, 4194304, 524288, 3670016, 4096, 512, 3584), 30000000: (4096, 4194304, 524288, 3670016, 2048, 256, 1792),
            60000000: (8192, 8388608, 1048576, 7340032, 2048, 256, 1792)}
    for sr, w in want.items():
        dec, tbw, d = O.geometry(sr)
        got = (dec, d.fft_size, d.overlap_length, d.input_size, d.fft_inv_size, d.scrap, d.post_input_size)
        assert got == w, (sr, got, w)
        assert d.post_decimation == 2


@needs_ref
def test_geometry_bitexact_vs_reference():
    for sr in RATES:
        dec = L.orc_fft_decimation_rate(sr, 5400)
        tbw = L.orc_relative_transition_bw(sr, 250)
        assert dec == R.compute_fft_decimation_rate(sr, 5400)
        assert tbw == R.compute_filter_relative_transition_bw(sr, 250)
        for fs in [0.0, -0.25776, 0.1234, 0.4999, -0.37, 1e-4, -0.4999]:
            d, rc = O.ddc_init(tbw, dec, fs)
            buf = C.create_string_buffer(R.ref_sizeof_fastddc())
            rc2 = R.fastddc_init(buf, tbw, dec, fs)
            iv = (C.c_int32 * 13)()
            fv = (C.c_float * 5)()
            R.ref_fastddc_fields(buf, iv, fv)
            mine = [d.pre_decimation, d.post_decimation, d.taps_length, d.taps_min_length, d.overlap_length, d.fft_size,
                    d.fft_inv_size, d.input_size, d.post_input_size, d.startbin, d.v, d.offsetbin, d.scrap]
            assert mine == list(iv) and rc == rc2
            assert [d.pre_shift, d.post_shift, d.dsa_sindelta, d.dsa_cosdelta, d.dsa_rate] == list(fv)
    for x in [0, 1, 2, 3, 4, 7, 8, 46, 1000, 4097 * 4]:
        assert L.orc_next_pow2(x) == R.next_pow2(x)


@needs_ref
def test_taps_bitexact_vs_reference():
    for n, lo, hi in [(257, -0.3, 0.2), (4097, 0.25 - 1 / 64, 0.25 + 1 / 64), (32769, -0.11 - 1 / 512, -0.11 + 1 / 512)]:
        a = np.zeros(n, np.complex64)
        b = np.zeros(n, np.complex64)
        L.orc_bandpass_taps(a, n, lo, hi)
        R.firdes_bandpass_c(b, n, lo, hi, 2)    # WINDOW_HAMMING
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@needs_ref
@pytest.mark.parametrize("sr", [250000, 2000000])
def test_channelizer_full_fold_bitexact_vs_reference(sr):
    """fastddc_inv_cc of the reference itself (with the oracle FFT standing in for fftw3f) == oracle FOLD_FULL."""
    rng = np.random.default_rng(7)
    dec, tbw, d0 = O.geometry(sr)
    fs = L.orc_channel_shift_rate(sr, 10000000, 10063000)
    N, M = d0.fft_size, d0.fft_inv_size
    c1 = L.orc_channelizer_create(dec, tbw, fs, O.FOLD_FULL)
    c2 = R.fft_channelizer_create(dec, tbw, fs)
    c3 = L.orc_channelizer_create(dec, tbw, fs, O.FOLD_SLICE)
    for blk in range(3):
        X = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
        o1, o2, o3 = (np.zeros(M, np.complex64) for _ in range(3))
        n1 = L.orc_channelizer_execute(c1, X, o1)
        n2 = R.ref_channelizer_execute(c2, X.copy(), o2)
        n3 = L.orc_channelizer_execute(c3, X, o3)
        assert n1 == n2 == n3 == d0.post_input_size // 2
        assert np.array_equal(o1[:n1].view(np.uint32), o2[:n1].view(np.uint32))
        # pass-band slice vs the reference's all-bin fold: leakage bound on a white spectrum (DESIGN.md tolerance)
        assert np.linalg.norm(o3[:n1] - o1[:n1]) / np.linalg.norm(o1[:n1]) < 2e-4
    L.orc_channelizer_destroy(c1)
    L.orc_channelizer_destroy(c3)
    R.fft_channelizer_destroy(c2)


def test_crc_check_value_and_table():
    b = np.frombuffer(b"123456789", np.uint8).copy()
    assert L.orc_crc16(b, 9, 0xFFFF) ^ 0xFFFF == 0x906E        # CRC-16/X-25 check value
    z = np.array([1], np.uint8)
    assert L.orc_crc16(z, 1, 0) == 0x1189                       # crc.c:8 table[1]
    z = np.array([255], np.uint8)
    assert L.orc_crc16(z, 1, 0) == 0x0F78                       # crc.c table[255]


@needs_ref
def test_crc_vs_reference():
    rng = np.random.default_rng(3)
    for n in [1, 2, 8, 66, 945]:
        b = rng.integers(0, 256, n, dtype=np.uint8)
        assert L.orc_crc16(b, n, 0xFFFF) == R.crc16_ccitt(b, n, 0xFFFF)


@needs_ref
def test_viterbi_bitexact_vs_reference():
    rng = np.random.default_rng(4)
    for nbits in [540, 1080, 1260, 2160, 2520, 3240, 5040, 7560]:
        for trial in range(2):
            syms = rng.integers(0, 256, 2 * nbits, dtype=np.uint8)
            if trial == 1:      # a valid codeword with noise
                bits = rng.integers(0, 2, nbits, dtype=np.uint8)
                chips = np.zeros(2 * nbits, np.uint8)
                L.orc_conv_encode27(bits, nbits, chips)
                syms = np.clip(chips.astype(np.int32) * 255 + rng.normal(0, 60, 2 * nbits), 0, 255).astype(np.uint8)
            o1 = np.zeros((nbits + 7) // 8, np.uint8)
            o2 = np.zeros((nbits + 7) // 8, np.uint8)
            L.orc_viterbi27(syms, nbits, o1)
            v = R.create_viterbi27(nbits)
            R.init_viterbi27(v, 0)
            R.update_viterbi27_blk(v, syms, nbits)
            R.chainback_viterbi27(v, o2, nbits, 0)
            R.delete_viterbi27(v)
            assert np.array_equal(o1, o2)


def test_viterbi_roundtrip_with_errors():
    rng = np.random.default_rng(5)
    nbits = 1080
    bits = rng.integers(0, 2, nbits, dtype=np.uint8)
    bits[-6:] = 0
    chips = np.zeros(2 * nbits, np.uint8)
    L.orc_conv_encode27(bits, nbits, chips)
    syms = (chips * 255).astype(np.uint8)
    for p in [100, 700, 1500]:
        syms[p] ^= 255
    out = np.zeros(nbits // 8, np.uint8)
    L.orc_viterbi27(syms, nbits, out)
    assert np.array_equal(np.unpackbits(out), bits)     # chainback packs MSB first


def test_known_sequences_from_reference_text():
    # A (first 127 bits of hfdl.c:420-437) and the M1 base sequence (hfdl.c:441-447) are m-sequences:
    # 64 ones, cyclic autocorrelation -1 off-peak
    A = np.unpackbits(np.array(list((C.c_uint8 * 16).in_dll(L, "orc_A_octets")), np.uint8))[:127]
    M = np.array(list((C.c_uint8 * 127).in_dll(L, "orc_M1_bits")), np.uint8)
    for s in (A, M):
        assert s.sum() == 64
        pm = 1 - 2 * s.astype(int)
        for k in range(1, 127):
            assert (pm * np.roll(pm, k)).sum() == -1
    # scrambler: ARINC 635 x^15+x+1, preset 110100101011001; first 32 bits recorded in SURVEY appendix A
    out = np.zeros(240, np.uint8)
    L.orc_scrambler_bits(out, 240)
    assert "".join(map(str, out[:32])) == "01100011001000110111101110000100"
    assert np.array_equal(out[:120], out[120:]) and out[:120].sum() == 49
    # mode table / PDU sizes (hfdl.c:81-138)
    assert [L.orc_pdu_len_octets(i) for i in range(8)] == [68, 135, 270, 405, 158, 315, 630, 945]


def test_fcs_rules():
    for M1 in range(8):
        for kind in (0, 1):
            p = np.frombuffer(O.make_pdu(M1, kind, seed=M1 * 3 + kind), np.uint8).copy()
            assert L.orc_pdu_crc_good(p, p.size) == 1
            p[3] ^= 0x10
            assert L.orc_pdu_crc_good(p, p.size) == 0


def test_timing_build_fft_standin_matches_numpy():
    """oracle/_ref/libref_fast.so (the timed CPU baseline only): when fftw3f is not on the machine its large transforms run in
    oracle/ref_shim/fft4step.c (cache-blocked four-step FFT on a worker pool) -- same unnormalised DFT, both signs, in place too"""
    R = O.reflib_fast()
    if R is None or not hasattr(R, "ref_fft_run"):
        pytest.skip("oracle/_ref/libref_fast.so not built")
    cfp = np.ctypeslib.ndpointer(np.complex64, flags="C")
    R.ref_fft_run.argtypes = [cfp, cfp, C.c_int32, C.c_int32]
    R.csdr_fft_init(4)
    rng = np.random.default_rng(8)
    for lg in (11, 12, 18, 19, 20, 22):
        n = 1 << lg
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        y = np.zeros(n, np.complex64)
        for fwd in (1, 0):
            R.ref_fft_run(x, y, n, fwd)
            ref = np.fft.fft(x.astype(np.complex128)) if fwd else np.fft.ifft(x.astype(np.complex128)) * n
            assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 5e-7, (lg, fwd)
        R.ref_fft_run(x, y, n, 1)
        z = x.copy()
        R.ref_fft_run(z, z, n, 1)                       # in place == out of place
        assert np.array_equal(z.view(np.uint32), y.view(np.uint32))

