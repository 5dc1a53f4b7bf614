"""Parity cases shared by the GPU tests (real libhfdl_b200.so, -m gpu) and the CPU logic tests
(same kernels compiled for host emulation, tests/cusim).  Every case drives the C ABI of
include/hfdl_b200.h and compares with the CPU oracle on the same seeded input."""
import ctypes as C

import numpy as np

import orclib as O
from pdu_forge import Forge
import dumphfdl_b200.api as A

# float tolerances (relative L2).  The reference itself is built -ffast-math, so float I/Q is
# compared within tolerance and only PDU octets / integer stages bit-exactly.
TOL_FFT = 2e-6          # forward spectrum vs float64 FFT
TOL_DDC = 2e-5          # channeliser output vs oracle (slice fold); closed-form phase vs the reference's recursion
TOL_DEMOD = 5e-4        # AGC / MF / EQ checkpoints (feedback loops amplify 1-ulp libm differences)
MEASURED = []           # float-checkpoint errors of the case_frontend runs of this process (tests may dump them)

CF = 10000000


def rel(a, b):
    n = min(a.size, b.size)
    return float(np.linalg.norm(a[:n] - b[:n]) / max(np.linalg.norm(b[:n]), 1e-30))


def case_fft(lib, sizes, batch=2, seed=1):
    rng = np.random.default_rng(seed)
    for n in sizes:
        x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
        y = A.fft_forward(x, lib=lib)
        ref = np.fft.fft(x.astype(np.complex128), axis=1)
        assert rel(y.ravel(), ref.ravel()) < TOL_FFT, n
        o = np.zeros(n, np.complex64)
        O.lib().orc_fft(x[0].copy(), o, n, 1)
        assert rel(y[0], o) < TOL_FFT
    # linearity + a pure tone lands in one bin (size-independent properties)
    n = sizes[-1]
    k = 5 * n // 16 + 3
    t = np.exp(2j * np.pi * k * np.arange(n) / n).astype(np.complex64)
    y = A.fft_forward(t, lib=lib)[0] if False else A.fft_forward(t, lib=lib).ravel()
    assert abs(abs(y[k]) - n) / n < 1e-4
    y[k] = 0
    assert np.abs(y).max() / n < 1e-4


def case_viterbi(lib, sizes, seed=4, frames=3):
    rng = np.random.default_rng(seed)
    L = O.lib()
    for nbits in sizes:
        syms = rng.integers(0, 256, (frames, 2 * nbits), dtype=np.uint8)
        bits = rng.integers(0, 2, nbits, dtype=np.uint8)
        chips = np.zeros(2 * nbits, np.uint8)
        L.orc_conv_encode27(bits, nbits, chips)
        syms[0] = np.clip(chips.astype(np.int32) * 255 + rng.normal(0, 70, 2 * nbits), 0, 255).astype(np.uint8)
        got = A.viterbi27(syms, nbits, lib=lib)
        for i in range(frames):
            want = np.zeros((nbits + 7) // 8, np.uint8)
            L.orc_viterbi27(syms[i].copy(), nbits, want)
            assert np.array_equal(got[i], want), (nbits, i)      # bit-exact


def case_fec(lib, modes, seed=2, noise=0.25):
    rng = np.random.default_rng(seed)
    L = O.lib()
    for M1 in modes:
        pdus = [O.make_pdu(M1, k % 2, seed=seed * 100 + M1 * 10 + k) for k in range(3)]
        nsym = [2160, 5040][M1 // 4]
        S = np.zeros((3, nsym), np.complex64)
        for k, p in enumerate(pdus):
            sym = np.zeros(5040, np.complex64)
            n = L.orc_encode_user_data(np.frombuffer(p, np.uint8).copy(), M1, sym)
            assert n == nsym
            nz = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (noise if k < 2 else 0.7)
            S[k] = ((sym[:n] + nz) * np.exp(1j * 0.05)).astype(np.complex64)
        for bm in (0, 0xFFFFFFFF):
            X = (S * (-1 if bm else 1)).astype(np.complex64)
            out, crc, soft = A.fec_decode(X, M1, bm, want_soft=True, lib=lib)
            nenc = nsym * [1, 1, 2, 3][M1 % 4]
            for k in range(3):
                ro = np.zeros(948, np.uint8)
                rs = np.zeros(15120, np.uint8)
                ln = L.orc_decode_user_data(X[k].copy(), M1, bm, ro, rs.ctypes.data)
                assert np.array_equal(soft[k][:nenc], rs[:nenc]), (M1, k)       # soft bits bit-exact
                assert np.array_equal(out[k], ro[:ln]), (M1, k)                  # PDU octets bit-exact
                assert crc[k] == L.orc_pdu_crc_good(ro, ln)
            assert bytes(out[0]) == pdus[0] and crc[0] == 1                      # and equal to what was sent


def make_capture(sr, freqs, modes, dur, esn0=20.0, seed=3, amp=None, starts=None):
    amp = amp or 0.5 / max(2.0, np.sqrt(len(freqs)) * 2)
    rng = np.random.default_rng(seed)
    frames, truth = [], []
    for i, (f, m) in enumerate(zip(freqs, modes)):
        pdu = O.make_pdu(m, (i + m) % 2, seed=seed * 1000 + i)
        st = starts[i] if starts else 0.15 + 0.03 * i
        frames.append(O.tx_frame(f, m, st, pdu, cfo_hz=float(rng.uniform(-15, 15)), phase0=float(rng.uniform(0, 6.28)), amplitude=amp))
        truth.append((f, pdu))
    x = O.render(int(sr * dur), sr, CF, frames, noise_sigma=O.noise_sigma(amp, sr, esn0), seed=seed)
    return x, truth


def hostile_capture(name, sr=250000, f=10063000):
    """Captures the domain's edge cases are made of (one channel at f, centre CF): -> (complex64 capture, note).  What the
    reference does with them is whatever its demodulator does; the oracle must do exactly the same (test_oracle_hfdl_ref.py)
    and the kernels must follow the oracle (host emulation: test_cusim_logic.py)."""
    pd = [O.make_pdu(1, 0, 501), O.make_pdu(2, 3, 502), O.make_pdu(3, 1, 503)]
    amp, dur, esn0, seed = 0.1, 3.4, 20.0, 500
    post = None
    if name == "collision":            # two transmitters in one slot on one channel, 6 dB apart, different carrier offsets
        frames = [O.tx_frame(f, 1, 0.25, pd[0], cfo_hz=6.0, phase0=0.3, amplitude=0.1),
                  O.tx_frame(f, 2, 0.55, pd[1], cfo_hz=-9.0, phase0=2.0, amplitude=0.05)]
    elif name == "equal_power_collision":
        frames = [O.tx_frame(f, 1, 0.25, pd[0], cfo_hz=6.0, phase0=0.3, amplitude=0.07),
                  O.tx_frame(f, 1, 0.40, pd[1][: len(pd[0])] if len(pd[1]) >= len(pd[0]) else pd[0], cfo_hz=-4.0, phase0=1.0, amplitude=0.07)]
    elif name == "cfo_plus_70":        # carrier far off the dial frequency (Costas pull-in range)
        frames = [O.tx_frame(f, 1, 0.25, pd[0], cfo_hz=70.0, phase0=0.3, amplitude=amp)]
    elif name == "cfo_minus_45":
        frames = [O.tx_frame(f, 3, 0.25, pd[2], cfo_hz=-45.0, phase0=1.3, amplitude=amp)]
    elif name == "clipped":            # ADC overload: the capture is clipped at 60 % of the frame's peak before quantisation
        frames = [O.tx_frame(f, 2, 0.25, pd[1], cfo_hz=5.0, phase0=0.3, amplitude=0.5)]
        post = "clip"
    elif name == "cut_at_end":         # the capture ends in the middle of the data part
        frames = [O.tx_frame(f, 1, 0.25, pd[0], cfo_hz=5.0, phase0=0.3, amplitude=amp)]
        dur = 1.6
    elif name == "starts_mid_frame":   # the capture begins inside a frame's preamble; a complete frame follows
        frames = [O.tx_frame(f, 1, -0.2, pd[0], cfo_hz=5.0, phase0=0.3, amplitude=amp),
                  O.tx_frame(f, 3, 2.5, pd[2], cfo_hz=5.0, phase0=0.3, amplitude=amp)]
        dur = 5.3
    elif name == "adjacent_interferer":   # a 30 dB stronger carrier 1.5 kHz outside the channel's pass band
        frames = [O.tx_frame(f, 2, 0.25, pd[1], cfo_hz=-3.0, phase0=0.3, amplitude=0.02)]
        post = "interferer"
        esn0 = 25.0
        amp = 0.02
    elif name == "dc_and_weak":        # a DC spur of the SDR at the centre frequency and a frame 40 dB below full scale
        frames = [O.tx_frame(f, 1, 0.25, pd[0], cfo_hz=2.0, phase0=0.3, amplitude=0.01)]
        post = "dc"
        amp = 0.01
    else:
        raise KeyError(name)
    x = O.render(int(sr * dur), sr, CF, frames, noise_sigma=O.noise_sigma(amp, sr, esn0), seed=seed)
    if post == "clip":
        lim = np.float32(0.3)
        x = (np.clip(x.real, -lim, lim) + 1j * np.clip(x.imag, -lim, lim)).astype(np.complex64)
    elif post == "interferer":
        t = np.arange(x.size) / sr
        x = (x + 0.6 * np.exp(2j * np.pi * (f - CF + 1440 + 2900.0) * t)).astype(np.complex64)
    elif post == "dc":
        x = (x + np.complex64(0.2 + 0.1j)).astype(np.complex64)
    return x


def random_scenario(seed, sr=250000):
    """A seeded random job: 1-3 channels, up to three frames per channel back to back or overlapping the next slot, modes,
    PDU kinds, carrier offsets (+-30 Hz), amplitudes (26 dB range) and Es/N0 (3-25 dB) drawn at random -> (freqs, capture)"""
    rng = np.random.default_rng(9000 + seed)
    nch = int(rng.integers(1, 4))
    freqs = sorted(rng.choice(np.arange(9890, 10111, 13), nch, replace=False) * 1000)
    freqs = [int(f) for f in freqs]
    frames, t_end = [], 0.0
    amp0 = float(10 ** rng.uniform(-2.0, -0.7))
    for f in freqs:
        t = float(rng.uniform(0.05, 0.6))
        for _ in range(int(rng.integers(1, 4))):
            m = int(rng.integers(0, 8))
            pdu = O.make_pdu(m, int(rng.integers(0, 5)), int(rng.integers(1, 1 << 30)))
            frames.append(O.tx_frame(f, m, t, pdu, cfo_hz=float(rng.uniform(-30, 30)), phase0=float(rng.uniform(0, 6.28)),
                                     amplitude=amp0 * float(10 ** rng.uniform(-0.3, 0.3))))
            t += (2.4615 if m < 4 else 4.923) + float(rng.choice([0.0, 0.0, 0.35, -0.4]))      # next slot, a gap, or an overlap
        t_end = max(t_end, t)
    esn0 = float(rng.uniform(3, 25))
    x = O.render(int(sr * (t_end + 0.4)), sr, CF, frames, noise_sigma=O.noise_sigma(amp0, sr, esn0), seed=seed)
    return freqs, x


HOSTILE = ("collision", "equal_power_collision", "cfo_plus_70", "cfo_minus_45", "clipped", "cut_at_end", "starts_mid_frame",
           "adjacent_interferer", "dc_and_weak")


def case_hostile(lib, name, sr=250000, f=10063000, batch=4):
    """the kernels on a hostile capture: PDUs (possibly none, possibly with bit errors), frame positions, demodulator and
    frame counters and the continuous float checkpoints equal the oracle's"""
    x = hostile_capture(name, sr, f)
    p = run_oracle(sr, [f], x, A.SFMT_CF32, 0, ["agc", "mf", "eq"])
    ref = p.pdus()
    fe = A.Frontend(sr, CF, [f], max_blocks_per_batch=batch, capture_channel=0, capture_max=1 << 20, lib=lib)
    fe.push(x)
    fe.flush()
    got = fe.pdus()
    compare_pdus(got, ref)
    assert fe.stats(0) == p.stats(0)
    check_counters(fe, p, [f], ref)
    check_front(got)
    for tap in ("agc", "mf", "eq"):
        a, b = fe.checkpoint(tap), p.capture(0, tap)
        assert a.size == b.size and rel(a, b) < TOL_DEMOD, (name, tap, rel(a, b))
    fe.close()
    return len(got)


def run_oracle(sr, freqs, raw, sfmt, capture_ch=None, taps=()):
    p = O.Pipeline(sr, CF, freqs, fold_mode=O.FOLD_SLICE, nthreads=8)
    if capture_ch is not None:
        p.set_capture(capture_ch, list(taps), 1 << 22)
    p.feed(raw, sfmt)
    return p


def compare_pdus(got, ref, truth=None, subset=False):
    g = sorted((q.freq, q.sample_cnt_end, q.data(), q.M1, q.crc_good, q.sample_cnt_a2) for q in got)
    r = sorted((q.freq, q.sample_cnt_end, q.data(), q.M1, q.crc_good, q.sample_cnt_a2) for q in ref)
    assert g == r, "PDU list differs from the oracle's"
    if truth is not None and subset:
        # the demodulator (reference and oracle alike) may miss a frame that follows another one closely; what it
        # does deliver with a good FCS must be a transmitted PDU
        assert all((f, d) in set(truth) for f, _, d, _, crc, _ in g if crc)
    elif truth is not None:
        assert sorted((f, d) for f, _, d, _, _, _ in g) == sorted(truth)


def check_counters(fe, p, freqs, ref):
    """hfdl_b200_channel_counters (the statsd metrics of doc/STATSD_METRICS.md) vs the oracle: preamble counters from the
    demodulator, frame / LPDU counters from the front parser restated in orc_pdu_front_parse"""
    for c, f in enumerate(freqs):
        k = fe.counters(c)
        a1, a2, m1, frames = p.stats(c)
        assert (k.freq, k.A1_found, k.A2_found, k.M1_found, k.M1_not_found) == (f, a1, a2, m1, p.m1_not_found(c))
        fr = [O.pdu_front(q.data()) for q in ref if q.freq == f]
        assert k.frames_processed == len(fr) == frames
        assert (k.frames_good, k.frames_bad_fcs, k.frames_too_short) == tuple(sum(1 for v in fr if v[0] == s) for s in (0, 1, 2))
        assert (k.frames_air2gnd, k.frames_gnd2air) == (sum(1 for v in fr if v[0] == 0 and v[1] == 1), sum(1 for v in fr if v[0] == 0 and v[1] == 0))
        assert (k.lpdus_processed, k.lpdus_good, k.lpdus_bad_fcs, k.lpdus_too_short) == tuple(sum(v[i] for v in fr) for i in (2, 3, 4, 5))
        assert abs(k.noise_floor - p.noise_floor(c)) / p.noise_floor(c) < 1e-3
        assert abs(fe.noise_floor(c) - p.noise_floor(c)) / p.noise_floor(c) < 1e-3


def check_front(got):
    for q in got:
        w = O.pdu_front(q.data())
        assert (q.frame_status, q.direction, q.lpdus_processed, q.lpdus_good, q.lpdus_bad_fcs, q.lpdus_too_short, q.lpdu_good_mask) == w
        assert q.crc_good == (w[0] == 0)


def case_front_parser(lib, nfuzz=300):
    pd = [O.make_pdu(m, k, 7 + m) for m in range(8) for k in range(5)] + [b"\x03", b"\x00" * 10, b"\x13\x05"]
    # constructed MPDUs / SPDUs, one per branch of mpdu.c:56-159 / lpdu.c:129-150 / spdu.c:55-70, and mutated ones (the same
    # generator is checked against the reference's own pdu_decoder_thread in tests/test_oracle_front_ref.py)
    F = Forge(np.random.default_rng(77))
    pd += F.branches() + F.fuzz(nfuzz)
    got = A.pdu_front_parse(pd, lib=lib)
    for p, g in zip(pd, got):
        w = O.pdu_front(p)
        assert tuple(g[:7]) == w and g[7] == (w[0] == 0), (len(p), bytes(p[:16]).hex(), g, w)
    # every branch is hit at least once: good / bad_fcs / too_short frames, both directions, bad and short LPDUs
    assert {g[0] for g in got} == {0, 1, 2} and {g[1] for g in got if g[0] == 0} == {0, 1}
    assert any(g[4] for g in got) and any(g[5] for g in got) and max(g[2] for g in got) == 30


def case_tapslice(lib, sr, freqs):
    """HFDL_B200_CP_TAPSLICE (fft_channelizer_create, fastddc.c:217-252, reduced to the M bins the slice fold uses) vs the
    oracle's tap-spectrum slice; device order = inverse-FFT input order = the oracle's slice with its halves swapped"""
    fe = A.Frontend(sr, CF, freqs, max_blocks_per_batch=2, lib=lib)
    for c, f in enumerate(freqs):
        t = fe.checkpoint("tapslice", c)
        o = np.fft.fftshift(O.slice_taps(sr, CF, f))
        assert t.size == o.size == fe.geom.fft_inv_size
        assert rel(t, o) < 1e-5, (f, rel(t, o))
    fe.close()


def case_frontend(lib, sr, freqs, modes, dur, sfmt=A.SFMT_CF32, batch=4, check_floats=True, ragged=False, seed=3, ragged_seed=9, esn0=20.0,
                  check_truth=True, starts=None, tol_ddc=TOL_DDC, tol_demod=TOL_DEMOD):
    x, truth = make_capture(sr, freqs, modes, dur, seed=seed, esn0=esn0, starts=starts)
    if sfmt == A.SFMT_CS16:
        raw = np.zeros(2 * x.size, np.int16)
        O.lib().orc_quantize_cs16(x, x.size, raw)
    elif sfmt == A.SFMT_CU8:
        raw = np.zeros(2 * x.size, np.uint8)
        O.lib().orc_quantize_cu8((x * 0.9).astype(np.complex64), x.size, raw)
    else:
        raw = x
    p = run_oracle(sr, freqs, raw, sfmt, 0, ["ddc", "agc", "mf", "eq"])
    ref = p.pdus()
    fe = A.Frontend(sr, CF, freqs, sample_format=sfmt, max_blocks_per_batch=batch, capture_channel=0, capture_max=1 << 18, lib=lib)
    g = fe.geom
    assert (g.fft_size, g.input_size, g.fft_inv_size, g.scrap) == (p.ddc.fft_size, p.ddc.input_size, p.ddc.fft_inv_size, p.ddc.scrap)
    if ragged:
        rng = np.random.default_rng(ragged_seed)
        per = raw.size // x.size
        i = 0
        while i < x.size:
            k = int(rng.integers(1, 3 * g.input_size))
            fe.push(raw[i * per:(i + k) * per])
            i += k
        fe.push(raw[:0])
    else:
        fe.push(raw)
    fe.flush()
    got = fe.pdus()
    compare_pdus(got, ref, truth if (sfmt != A.SFMT_CU8 and check_truth) else None)
    for c in range(len(freqs)):
        assert fe.stats(c) == p.stats(c)
    check_counters(fe, p, freqs, ref)
    check_front(got)
    if check_floats:
        # spectrum of the last processed block vs the oracle's last spectrum (oracle holds the swapped one)
        spec = fe.checkpoint("spectrum", -1)
        osp = np.fft.ifftshift(p.last_spectrum())
        assert rel(spec, osp) < TOL_FFT * 2, rel(spec, osp)
        ddc = fe.checkpoint("ddc", 0)
        od = p.capture(0, "ddc")
        assert rel(ddc, od[-ddc.size:]) < tol_ddc, rel(ddc, od[-ddc.size:])
        MEASURED.append(dict(sr=sr, nch=len(freqs), sfmt=sfmt, spectrum=rel(spec, osp), ddc=rel(ddc, od[-ddc.size:])))
        for name in ("agc", "mf", "eq"):
            a, b = fe.checkpoint(name), p.capture(0, name)
            if a.size == b.size:
                MEASURED[-1][name] = rel(a, b)
                if name == "eq":            # where the equaliser output differs most (loops amplify: a frame's edges / noise-only stretches)
                    e = np.abs(a - b)
                    MEASURED[-1]["eq_worst_symbol"] = int(np.argmax(e))
                    MEASURED[-1]["eq_rel_first_half"] = rel(a[: a.size // 2], b[: a.size // 2])
            assert a.size == b.size and rel(a, b) < tol_demod, (name, a.size, b.size, rel(a, b), MEASURED[-1])
        # metadata that goes into hfdl_pdu_metadata (hfdl.c:1061-1067)
        for q, r in zip(sorted(got, key=lambda z: (z.sample_cnt_end, z.freq)), sorted(ref, key=lambda z: (z.sample_cnt_end, z.freq))):
            assert abs(q.freq_err_hz - r.freq_err_hz) < 1e-2
            assert abs(q.signal_level - r.signal_level) / r.signal_level < 1e-3
            assert abs(q.noise_floor_lin - r.noise_floor) / r.noise_floor < 1e-3
            assert q.bit_rate == r.bit_rate and q.slot == r.slot
            assert (q.train_bits_bad, q.train_bits_total) == (r.train_bits_bad, r.train_bits_total)
    fe.close()
    return len(got)


def case_errors_and_empty_inputs(lib):
    """Error behaviour of the boundary and the degenerate inputs: what the reference rejects before it builds its blocks
    (main.c:214-226 check_frequency_span, main.c:699-706 / fastddc.c:46-80 geometry, input-common.h sample formats) is
    rejected by hfdl_b200_create; an empty push, a push shorter than one overlap-save block and a flush of an empty stream
    are valid and give no PDU (fft.c:38-55 simply waits for input_size samples)."""
    sr, f = 250000, 10063000
    for kw, args in (
            (dict(), (sr, CF, [])),                                    # no channel
            (dict(), (sr, CF, [CF + sr // 2])),                        # |centre - channel| >= sample_rate / 2 (main.c:217)
            (dict(), (sr, CF, [f, CF - sr // 2 - 1000])),              # ... any channel of the list
            (dict(sample_format=0), (sr, CF, [f])),                    # SFMT_UNDEF
            (dict(sample_format=4), (sr, CF, [f])),                    # beyond SFMT_CF32
            (dict(), (5000, CF, [CF + 1000])),                         # rate below the 5400 Hz the demodulator runs at: decimation < 1
    ):
        try:
            A.Frontend(*args, lib=lib, **kw)
        except RuntimeError:
            continue
        raise AssertionError("hfdl_b200_create accepted %r %r" % (args, kw))
    vp = C.c_void_p()
    assert lib.hfdl_b200_create(None, None) == -1 and lib.hfdl_b200_create(C.byref(vp), None) == -1
    fe = A.Frontend(sr, CF, [f, 9952000], max_blocks_per_batch=3, lib=lib)
    g = fe.geom
    assert fe.push(np.zeros(0, np.complex64)) == 0 and fe.flush() == 0 and fe.pdus() == []
    assert lib.hfdl_b200_push_samples(fe.h, None, 5) == -1 and lib.hfdl_b200_push_samples(fe.h, None, -1) == -1
    assert lib.hfdl_b200_push_samples(fe.h, None, 0) == 0
    fe.push(np.zeros(g.input_size - 1, np.complex64))                 # one sample short of a block: nothing to transform yet
    fe.flush()
    assert fe.pdus() == [] and fe.launches() == 0
    fe.push(np.zeros(1 + 2 * g.input_size, np.complex64))             # silence: blocks run, no preamble, no PDU
    fe.flush()
    assert fe.pdus() == [] and fe.launches() > 0
    for c in range(2):
        assert fe.stats(c) == (0, 0, 0, 0)
        k = fe.counters(c)
        assert k.freq == fe.freqs[c] and k.frames_processed == 0 and k.M1_not_found == 0 and k.A2_found == 0
    st = (C.c_int32 * 4)()
    assert lib.hfdl_b200_channel_stats(fe.h, 2, st) == -1 and lib.hfdl_b200_channel_stats(fe.h, -1, st) == -1
    assert lib.hfdl_b200_channel_counters(fe.h, 2, C.byref(A.Counters())) == -1
    lvl = C.c_float()
    assert lib.hfdl_b200_channel_noise_floor(fe.h, 7, C.byref(lvl)) == -1
    one = A.Pdu()
    assert lib.hfdl_b200_pop_pdu(fe.h, C.byref(one)) == 0             # empty queue: 0, not an error
    fe.close()


def case_pruned_spectrum(lib, sr, freqs, modes, dur, batch, seed=61):
    """Without a capture channel the last FFT pass stores only the spectrum granules some channel's slice reads
    (production mode); with one it stores every bin (parity / debug mode).  Both must give the oracle's PDUs."""
    x, truth = make_capture(sr, freqs, modes, dur, seed=seed)
    ref = run_oracle(sr, freqs, x, A.SFMT_CF32).pdus()
    for cap in (-1, 0):
        fe = A.Frontend(sr, CF, freqs, max_blocks_per_batch=batch, capture_channel=cap, capture_max=1024 if cap >= 0 else 0, lib=lib)
        fe.push(x)
        fe.flush()
        compare_pdus(fe.pdus(), ref, truth)
        if cap < 0:
            try:
                fe.checkpoint("spectrum", -1)
                assert False, "the pruned spectrum must not be handed out as a checkpoint"
            except RuntimeError:
                pass
        fe.close()
    return len(ref)


def case_frontend_stream(lib, sr, freqs, plan, dur, batch, esn0=20.0, seed=31, push_blocks=None, submit_poll=False):
    """Several frames per channel over many batches: `plan` = [(channel index, M1, start second), ...].
    Exercises the cross-batch pipeline (sub-range schedule, shared work arrays, deferred PDU collection):
    PDUs, counters and the continuous AGC / matched-filter / equaliser checkpoints must equal the oracle's."""
    amp = 0.5 / max(2.0, np.sqrt(len(freqs)) * 2)
    rng = np.random.default_rng(seed)
    frames, truth = [], []
    for i, (ch, m, st) in enumerate(plan):
        pdu = O.make_pdu(m, i % 2, seed=seed * 1000 + i)
        frames.append(O.tx_frame(freqs[ch], m, st, pdu, cfo_hz=float(rng.uniform(-15, 15)), phase0=float(rng.uniform(0, 6.28)), amplitude=amp))
        truth.append((freqs[ch], pdu))
    x = O.render(int(sr * dur), sr, CF, frames, noise_sigma=O.noise_sigma(amp, sr, esn0), seed=seed)
    p = run_oracle(sr, freqs, x, A.SFMT_CF32, 0, ["agc", "mf", "eq"])
    ref = p.pdus()
    fe = A.Frontend(sr, CF, freqs, max_blocks_per_batch=batch, capture_channel=0, capture_max=1 << 20, lib=lib)
    isz = fe.geom.input_size
    got = []
    if push_blocks:
        for i in range(0, x.size, push_blocks * isz):
            fe.push(x[i:i + push_blocks * isz])
            if submit_poll:                        # the block shim's use: queue what is there, pick up what has finished
                fe.submit()
                fe.poll()
            got += fe.pdus()                       # streaming use: PDUs are picked up as the pipeline delivers them
    else:
        fe.push(x)
    fe.flush()
    got += fe.pdus()
    compare_pdus(got, ref, truth, subset=True)
    for c in range(len(freqs)):
        assert fe.stats(c) == p.stats(c)
    check_counters(fe, p, freqs, ref)
    for name in ("agc", "mf", "eq"):
        a, b = fe.checkpoint(name), p.capture(0, name)
        assert a.size == b.size and rel(a, b) < TOL_DEMOD, name
    fe.close()
    return len(got)


def case_random_job(lib, seed, sr=250000, batch=5):
    """the kernels on a seeded random job (random_scenario): PDUs, positions, counters equal the oracle's"""
    freqs, x = random_scenario(seed, sr)
    p = run_oracle(sr, freqs, x, A.SFMT_CF32)
    ref = p.pdus()
    fe = A.Frontend(sr, CF, freqs, max_blocks_per_batch=batch, lib=lib)
    fe.push(x)
    fe.flush()
    got = fe.pdus()
    compare_pdus(got, ref)
    for c in range(len(freqs)):
        assert fe.stats(c) == p.stats(c)
    check_counters(fe, p, freqs, ref)
    check_front(got)
    fe.close()
    return len(got)


class HostMem:
    """'Device' memory of the host-emulation build: plain numpy buffers."""

    def upload(self, a):
        return np.ascontiguousarray(a).copy()

    def empty(self, nbytes):
        return np.zeros(nbytes, np.uint8)

    def ptr(self, h, byte_offset=0):
        return h.ctypes.data + byte_offset

    def copy(self, dst, dst_off, src, src_off, nbytes):
        import ctypes
        ctypes.memmove(dst.ctypes.data + dst_off, src.ctypes.data + src_off, nbytes)

    def sync(self):
        pass


class CudaMem:
    """Device memory through torch (GPU tests)."""

    def __init__(self):
        import torch
        self.t = torch

    def upload(self, a):
        a = np.ascontiguousarray(a)
        return self.t.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()

    def empty(self, nbytes):
        return self.t.zeros(nbytes, dtype=self.t.uint8, device="cuda")

    def ptr(self, h, byte_offset=0):
        return h.data_ptr() + byte_offset

    def copy(self, dst, dst_off, src, src_off, nbytes):
        dst[dst_off:dst_off + nbytes].copy_(src[src_off:src_off + nbytes])

    def sync(self):
        self.t.cuda.synchronize()


def case_sharded_spectrum(lib, mem, sr, freqs, modes, dur, nranks, batch, sfmt=A.SFMT_CF32, seed=71, starts=None, direct=False, cache=None):
    """Multi-GPU data path on ONE device: `nranks` frontends, rank r owning the channels freqs[r::nranks].  Per batch every
    rank transforms its share of the blocks for all channels (hfdl_b200_spectrum_slices), the slices change hands (here:
    plain copies standing in for the all-to-all) and every rank demodulates its channels (hfdl_b200_process_slices).
    direct = True: hfdl_b200_spectrum_slices_to -- the pack kernel stores into every rank's receive buffer itself.
    The PDUs must be those of one frontend that owns all channels (and the oracle's)."""
    assert batch % nranks == 0 and len(freqs) % nranks == 0
    x, truth = make_capture(sr, freqs, modes, dur, seed=seed, starts=starts)
    raw = x
    if sfmt == A.SFMT_CS16:
        raw = np.zeros(2 * x.size, np.int16)
        O.lib().orc_quantize_cs16(x, x.size, raw)
    bps = {A.SFMT_CS16: 4, A.SFMT_CF32: 8}[sfmt]
    fes = [A.Frontend(sr, CF, freqs[r::nranks], sample_format=sfmt, max_blocks_per_batch=batch, lib=lib) for r in range(nranks)]
    for fe in fes:
        fe.set_exchange(freqs, nranks)
    g = fes[0].geom
    isz, ovl, M = g.input_size, g.overlap_length, g.fft_inv_size
    cper = len(freqs) // nranks
    nb = x.size // isz
    nb -= nb % nranks                                     # every batch (the last, shorter one too) splits evenly
    rawb = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
    padded = np.concatenate([np.zeros(ovl * bps, np.uint8), rawb[: nb * isz * bps]])
    key = (sr, tuple(freqs), tuple(modes), dur, nranks, batch, sfmt, seed)
    if cache is not None and key in cache:              # a second variant of the same case: the single-frontend result is known
        want = cache[key]
    else:
        ref_fe = A.Frontend(sr, CF, freqs, sample_format=sfmt, max_blocks_per_batch=batch, lib=lib)
        ref_fe.push(rawb[: nb * isz * bps])
        ref_fe.flush()
        want = sorted((q.freq, q.sample_cnt_a2, q.sample_cnt_end, q.M1, q.crc_good, q.data()) for q in ref_fe.pdus())
        ref_fe.close()
        if cache is not None:
            cache[key] = want
    slice_bytes = M * 8
    done = 0
    keep = []                                             # buffers stay alive until the pipelines have drained
    while done < nb:
        B = min(batch, nb - done)
        bl = B // nranks
        sends = []
        recvs = [mem.empty(B * cper * slice_bytes) for _ in fes]
        for r, fe in enumerate(fes):
            first = done + r * bl
            buf = mem.upload(padded[first * isz * bps: (first * isz + ovl + bl * isz) * bps])
            keep.append(buf)
            if direct:
                fe.spectrum_slices_to(mem.ptr(buf), first, bl, [mem.ptr(h) for h in recvs], r * bl)
                continue
            send = mem.empty(nranks * bl * cper * slice_bytes)
            fe.spectrum_slices(mem.ptr(buf), first, bl, mem.ptr(send))
            sends.append(send)
        mem.sync()                                        # spectrum_slices ran on each frontend's own stream: device-wide sync
        part = bl * cper * slice_bytes
        for q, fe in enumerate(fes):
            if not direct:
                for r in range(nranks):
                    mem.copy(recvs[q], r * part, sends[r], q * part, part)
                mem.sync()
            fe.process_slices(mem.ptr(recvs[q]), B)
        keep += recvs
        keep += sends
        done += B
    got = []
    for fe in fes:
        fe.sync()
        got += fe.pdus()
        fe.close()
    got = sorted((q.freq, q.sample_cnt_a2, q.sample_cnt_end, q.M1, q.crc_good, q.data()) for q in got)
    assert got == want
    assert sorted((f, d) for f, _, _, _, _, d in got) == sorted(truth)
    return len(got)
