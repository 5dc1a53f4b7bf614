"""Properties of the oracle's liquid-dsp restatements (oracle/orc_liquid.c, the Kaiser designs and the resampler in
oracle/orc_dsp.c).  liquid-dsp itself is not available here ("parity unpinned"), so these are not parity
tests: they check each object against what its published definition implies, evaluated independently in float64 --
the designed taps against numpy's Bessel / sinc, the resampler's rate, pass-band gain and stop-band rejection, the AGC's
fixed point, the PSK maps and soft bits, the scrambler LFSR's period.  A misread formula fails here even if the oracle's
own transmitter and receiver agree with each other."""
import ctypes as C

import numpy as np
import pytest

import orclib as O


class CF(C.Structure):          # `float complex` by value == two packed floats in one SSE register (SysV x86-64)
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


def L():
    lib = O.lib()
    lib.orc_agc_init.argtypes = [C.c_void_p, C.c_float]
    lib.orc_agc_execute.argtypes = [C.c_void_p, CF]
    lib.orc_agc_execute.restype = CF
    lib.orc_psk_point.argtypes = [C.c_int, C.c_uint32]
    lib.orc_psk_point.restype = CF
    lib.orc_modem_demod.argtypes = [C.c_int, CF, C.c_void_p]
    lib.orc_modem_demod.restype = C.c_uint32
    lib.orc_modem_demod_soft.argtypes = [C.c_int, CF, C.c_void_p, C.c_void_p]
    lib.orc_msequence_init.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    lib.orc_msequence_advance.argtypes = [C.c_void_p]
    lib.orc_msequence_advance.restype = C.c_uint32
    return lib


def kaiser_beta(As):
    return 0.1102 * (As - 8.7) if As > 50 else 0.5842 * (As - 21) ** 0.4 + 0.07886 * (As - 21)


@pytest.mark.parametrize("n,fc,As,mu", [(289, 0.75 / 48, 40.0, 0.0),          # symsync prototype (symsync_crcf_create_kaiser(3, 3, *, 16))
                                        (15, 0.45, 40.0, 0.0),                # eqlms_cccf_create_lowpass(15, 0.45)
                                        (3585, 0.3822 / 256, 60.0, 0.0),      # resamp prototype at rate 0.6912 * 0.553 region
                                        (57, 0.2, 60.0, 0.3)])                # a fractional sample offset
def test_kaiser_design_matches_float64_evaluation(n, fc, As, mu):
    h = np.zeros(n, np.float32)
    O.lib().orc_firdes_kaiser(n, fc, As, mu, h)
    t = np.arange(n) - (n - 1) / 2 + mu
    beta = kaiser_beta(As)
    w = np.i0(beta * np.sqrt(1 - (2 * t / n) ** 2)) / np.i0(beta)       # liquid 1.3.x kaiser(): argument 2t/N
    ref = np.sinc(2 * fc * t) * w
    assert np.max(np.abs(h - ref)) < 2e-5
    # a low-pass with unity-ish pass band once normalised: the DC gain of sinc(2 fc t) is 1 / (2 fc)
    assert abs(h.sum() * 2 * fc - 1.0) < 0.02


@pytest.mark.parametrize("rate", [0.6912, 0.55296, 0.73728])                  # BASELINE configs 1/2, 3, 4/5
def test_resampler_rate_passband_and_stopband(rate):
    lib = O.lib()
    q = lib.orc_resamp_create(rate, 60.0)
    nx = 6000

    def run(f_in):
        qq = lib.orc_resamp_create(rate, 60.0)
        x = np.exp(2j * np.pi * f_in * np.arange(nx)).astype(np.complex64)
        y = np.zeros(nx + 8, np.complex64)
        ny = C.c_uint32()
        lib.orc_resamp_execute(qq, x, nx, y, C.byref(ny))
        lib.orc_resamp_destroy(qq)
        return y[: ny.value]

    y = run(0.05)
    assert abs(y.size - nx * rate) <= 1                                      # output / input = rate
    # Kaiser prototype of 2*7*256+1 taps, As = 60 dB: transition width (As - 7.95) / (14.36 * 3585) * 256 = 0.26 of the input rate
    # around the cut-off 0.515 * rate (msresamp's choice): flat below fc - 0.13, 60 dB down above fc + 0.13
    fc = min(0.515 * rate, 0.49)
    for f_in in (0.01, 0.05, 0.1, fc - 0.16):                                # pass band: a tone comes out at f / rate with unit gain
        y = run(f_in)[200:]
        k = np.arange(y.size)
        z = y * np.exp(-2j * np.pi * (f_in / rate) * k)
        assert abs(np.abs(z.mean()) - 1.0) < 0.01, (rate, f_in, np.abs(z.mean()))
        assert np.std(np.abs(y)) < 0.01
    y = run(fc)[200:]                                                        # the cut-off itself: -6 dB
    assert abs(20 * np.log10(np.sqrt(np.mean(np.abs(y) ** 2))) + 6.0) < 0.5
    for f_in in (fc + 0.14, 0.5):                                            # stop band (where the input rate leaves room for one)
        if f_in > 0.5 or f_in < fc + 0.135:
            continue
        y = run(f_in)[200:]
        assert 20 * np.log10(np.sqrt(np.mean(np.abs(y) ** 2)) + 1e-12) < -55, (rate, f_in)
    lib.orc_resamp_destroy(q)


@pytest.mark.parametrize("amp", [1e-3, 0.1, 10.0])
def test_agc_fixed_point_and_level(amp):
    lib = L()
    st = C.create_string_buffer(64)
    lib.orc_agc_init(st, 0.01)
    y = None
    for i in range(4000):
        ph = 0.37 * i
        y = lib.orc_agc_execute(st, CF(amp * np.cos(ph), amp * np.sin(ph)))
    g = np.frombuffer(st, np.float32, 3)[0]
    assert abs(np.hypot(y.re, y.im) - 1.0) < 0.01                            # unit output energy
    assert abs(1.0 / g - amp) / amp < 0.01                                   # agc_crcf_get_signal_level


@pytest.mark.parametrize("m", [1, 2, 3])
def test_psk_maps_and_soft_bits(m):
    lib = L()
    M = 1 << m
    st = C.create_string_buffer(64)
    pts = []
    for s in range(M):
        p = lib.orc_psk_point(m, s)
        z = complex(p.re, p.im)
        assert abs(abs(z) - 1) < 1e-6
        pts.append(z)
        assert lib.orc_modem_demod(m, CF(p.re * 0.7, p.im * 0.7), st) == s    # decision regions are cones: scale-invariant
        soft = (C.c_uint8 * 3)()
        lib.orc_modem_demod_soft(m, p, st, soft)
        assert [int(v > 127) for v in soft[:m]] == [(s >> (m - 1 - i)) & 1 for i in range(m)]      # MSB first, 255 = one
    # Gray map: neighbours on the circle differ in exactly one bit
    order = sorted(range(M), key=lambda s: np.angle(pts[s]) % (2 * np.pi))
    for a, b in zip(order, order[1:] + order[:1]):
        assert bin(a ^ b).count("1") == 1
    if m == 3:      # soft bit of the bit two neighbours disagree on is monotonic in the angle between them
        a, b = order[0], order[1]
        k = (a ^ b).bit_length() - 1
        vals = []
        for t in np.linspace(0.05, 0.95, 10):
            ang = np.angle(pts[a]) + t * (np.pi / 4)
            soft = (C.c_uint8 * 3)()
            lib.orc_modem_demod_soft(m, CF(np.cos(ang), np.sin(ang)), st, soft)
            vals.append(soft[m - 1 - k])
        d = np.diff(np.array(vals, int))
        assert (d >= 0).all() or (d <= 0).all()
        assert abs(vals[0] - vals[-1]) > 100


def test_scrambler_lfsr_is_maximal_length_in_both_conventions():
    """x^15 + x + 1 is primitive: with the arguments hfdl.c:333-345 passes for either liquid generation the register must
    run through all 2^15 - 1 states (the frame restarts it every 120 symbols, so nothing else tests this)"""
    lib = L()
    for conv, poly, init in ((0, 0x8002, 0x6959), (1, 0x4001, 0x4d4b)):
        st = C.create_string_buffer(64)
        lib.orc_msequence_init(st, 15, poly, init, conv)
        bits = np.array([lib.orc_msequence_advance(st) for _ in range(2 * 32767)], np.uint8)
        assert bits[:32767].sum() == 16384                                    # balance property of an m-sequence
        assert np.array_equal(bits[:32767], bits[32767:])                     # period divides 2^15 - 1 = 7 * 31 * 151 ...
        for p in (32767 // 7, 32767 // 31, 32767 // 151):                     # ... and none of its maximal proper divisors
            assert not np.array_equal(bits[:p], bits[p:2 * p])
