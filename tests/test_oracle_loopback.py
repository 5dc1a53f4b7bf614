"""Generator truth for the oracle: PDUs produced by the HFDL transmitter (oracle/orc_tx.c) must come
back octet-for-octet, FCS good, through the whole restated receive path (fft.c -> fastddc.c -> hfdl.c ->
viterbi27 -> crc) for all 8 M1 modes, raw sample formats, noise, and the pass-band-slice fold."""
import numpy as np
import pytest

import orclib as O

SR, CF, F = 250000, 10000000, 10063000


def capture(M1, kind, seed, esn0=20.0, cfo=7.0, dur=None, amp=0.1):
    pdu = O.make_pdu(M1, kind=kind, seed=seed)
    dur = dur or (5.6 if M1 >= 4 else 3.2)
    fr = O.tx_frame(F, M1, 0.25, pdu, cfo_hz=cfo, phase0=1.0, amplitude=amp)
    return pdu, O.render(int(SR * dur), SR, CF, [fr], noise_sigma=O.noise_sigma(amp, SR, esn0), seed=seed)


@pytest.mark.parametrize("M1", range(8))
def test_all_modes_decode_exact(M1):
    pdu, x = capture(M1, M1 % 2, 5 + M1)
    p = O.Pipeline(SR, CF, [F], fold_mode=O.FOLD_FULL, nthreads=2)
    p.feed(x)
    got = p.pdus()
    assert len(got) == 1
    q = got[0]
    assert q.data() == pdu and q.crc_good == 1 and q.M1 == M1
    assert q.slot == (b"S" if M1 < 4 else b"D")
    assert q.bit_rate == [300, 600, 1200, 1800][M1 % 4]          # hfdl.c:1072-1073
    assert p.stats(0) == (1, 1, 1, 1)


def test_slice_fold_same_pdus_and_close_floats():
    pdu, x = capture(3, 0, 11, esn0=25.0)
    outs = []
    for mode in (O.FOLD_FULL, O.FOLD_SLICE):
        p = O.Pipeline(SR, CF, [F], fold_mode=mode, nthreads=2)
        p.set_capture(0, ["ddc"], 1 << 16)
        p.feed(x)
        got = p.pdus()
        assert len(got) == 1 and got[0].data() == pdu
        outs.append(p.capture(0, "ddc"))
    a, b = outs
    assert a.size == b.size and a.size > 0
    # stated tolerance for dropping the aliases outside the pass-band slice
    assert np.linalg.norm(a - b) / np.linalg.norm(a) < 1e-3


def test_cs16_and_cu8_inputs():
    pdu, x = capture(1, 0, 21, esn0=25.0, amp=0.2)
    n = x.size
    s16 = np.zeros(2 * n, np.int16)
    O.lib().orc_quantize_cs16(x, n, s16)
    p = O.Pipeline(SR, CF, [F], nthreads=2)
    p.feed(s16, O.SFMT_CS16)
    assert [q.data() for q in p.pdus()] == [pdu]
    u8 = np.zeros(2 * n, np.uint8)
    O.lib().orc_quantize_cu8(x, n, u8)
    p = O.Pipeline(SR, CF, [F], nthreads=2)
    p.feed(u8, O.SFMT_CU8)
    assert [q.data() for q in p.pdus()] == [pdu]


def test_feed_in_ragged_pieces_equals_one_shot():
    pdu, x = capture(2, 1, 31)
    p1 = O.Pipeline(SR, CF, [F], nthreads=1)
    p1.set_capture(0, ["eq"], 1 << 16)
    p1.feed(x)
    p2 = O.Pipeline(SR, CF, [F], nthreads=1)
    p2.set_capture(0, ["eq"], 1 << 16)
    rng = np.random.default_rng(0)
    i = 0
    while i < x.size:
        k = int(rng.integers(1, 70000))
        p2.feed(x[i:i + k])
        i += k
    assert p2.feed(np.zeros(0, np.complex64)) == 0                   # empty input
    assert [q.data() for q in p1.pdus()] == [q.data() for q in p2.pdus()] == [pdu]
    assert np.array_equal(p1.capture(0, "eq"), p2.capture(0, "eq"))


def test_noise_only_and_silence_give_no_pdus():
    p = O.Pipeline(SR, CF, [F], nthreads=2)
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(SR * 2) + 1j * rng.standard_normal(SR * 2)).astype(np.complex64) * 0.01
    p.feed(x)
    p.feed(np.zeros(SR, np.complex64))
    assert p.pdus() == []


def test_two_channels_back_to_back_frames_and_low_snr():
    f2 = 10021000
    pd = [O.make_pdu(1, 0, 41), O.make_pdu(3, 1, 42), O.make_pdu(0, 0, 43)]
    frames = [O.tx_frame(F, 1, 0.2, pd[0], cfo_hz=-12, amplitude=0.05),
              O.tx_frame(F, 3, 0.2 + 2.4615, pd[1], cfo_hz=9, phase0=2, amplitude=0.05),
              O.tx_frame(f2, 0, 0.9, pd[2], cfo_hz=3, amplitude=0.08)]
    x = O.render(int(SR * 5.6), SR, CF, frames, noise_sigma=O.noise_sigma(0.05, SR, 15.0), seed=9)
    p = O.Pipeline(SR, CF, [F, f2], nthreads=2)
    p.feed(x)
    got = {(q.freq, q.data()) for q in p.pdus()}
    assert got == {(F, pd[0]), (F, pd[1]), (f2, pd[2])}
    assert all(q.crc_good for q in p.pdus())
