"""include/hfdl_b200_ring.h -- the spectrum ring that replaces the reference's barrier pair + single shared buffer
(block.c:35-43,90-120, fft.c:57-61, hfdl.c:663-664) on the CPU path.  Host C in the product library; exercised here
(a) through its C ABI with real threads and (b) under the oracle's demodulator: a pipeline whose channel workers read
spectra from the ring must produce exactly the PDUs and checkpoints of the barrier-style pipeline."""
import ctypes as C
import os
import threading
import time

import numpy as np
import pytest

import orclib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dumphfdl_b200", "libhfdl_b200.so")


def ring_lib():
    L = C.CDLL(LIB if os.path.exists(LIB) else os.path.join(ROOT, "oracle", "liboracle.so"))
    L.hfdl_spectrum_ring_create.restype = C.c_void_p
    L.hfdl_spectrum_ring_create.argtypes = [C.c_size_t, C.c_int32, C.c_int32]
    L.hfdl_spectrum_ring_destroy.argtypes = [C.c_void_p]
    L.hfdl_spectrum_ring_produce_begin.restype = C.POINTER(C.c_float)
    L.hfdl_spectrum_ring_produce_begin.argtypes = [C.c_void_p]
    L.hfdl_spectrum_ring_produce_end.argtypes = [C.c_void_p]
    L.hfdl_spectrum_ring_consume_begin.restype = C.POINTER(C.c_float)
    L.hfdl_spectrum_ring_consume_begin.argtypes = [C.c_void_p, C.c_int32]
    L.hfdl_spectrum_ring_consume_end.argtypes = [C.c_void_p, C.c_int32]
    L.hfdl_spectrum_ring_shutdown.argtypes = [C.c_void_p]
    L.hfdl_spectrum_ring_drain.argtypes = [C.c_void_p]
    L.hfdl_spectrum_ring_produced.restype = C.c_int64
    L.hfdl_spectrum_ring_produced.argtypes = [C.c_void_p]
    return L


@pytest.mark.parametrize("depth,consumers", [(1, 1), (1, 3), (2, 3), (4, 2)])
def test_every_consumer_sees_every_spectrum_once_in_order(depth, consumers):
    L = ring_lib()
    bins, nblocks = 64, 200
    r = L.hfdl_spectrum_ring_create(bins, depth, consumers)
    assert r
    seen = [[] for _ in range(consumers)]
    maxlead = [0]

    def consumer(c):
        rng = np.random.default_rng(c)
        while True:
            p = L.hfdl_spectrum_ring_consume_begin(r, c)
            if not p:
                break
            a = np.ctypeslib.as_array(p, shape=(2 * bins,))
            assert (a == a[0]).all()                       # a slot is never overwritten while it is being read
            seen[c].append(int(a[0]))
            if rng.random() < 0.05:
                time.sleep(0.001)
            L.hfdl_spectrum_ring_consume_end(r, c)

    th = [threading.Thread(target=consumer, args=(c,)) for c in range(consumers)]
    for t in th:
        t.start()
    for k in range(nblocks):
        p = L.hfdl_spectrum_ring_produce_begin(r)
        np.ctypeslib.as_array(p, shape=(2 * bins,))[:] = k
        L.hfdl_spectrum_ring_produce_end(r)
        maxlead[0] = max(maxlead[0], k + 1 - min(len(s) for s in seen))
    L.hfdl_spectrum_ring_drain(r)
    assert L.hfdl_spectrum_ring_produced(r) == nblocks
    L.hfdl_spectrum_ring_shutdown(r)
    for t in th:
        t.join(timeout=10)
        assert not t.is_alive()
    assert all(s == list(range(nblocks)) for s in seen)
    assert maxlead[0] <= depth + 1                         # the producer never runs more than `depth` spectra ahead
    L.hfdl_spectrum_ring_destroy(r)


def test_bad_arguments():
    L = ring_lib()
    assert not L.hfdl_spectrum_ring_create(0, 2, 1)
    assert not L.hfdl_spectrum_ring_create(16, 0, 1)
    assert not L.hfdl_spectrum_ring_create(16, 2, 0)


def test_ring_pipeline_equals_barrier_pipeline():
    sr, cf = 250000, 10000000
    freqs = [10063000, 9952000, 10101000]
    pd = [O.make_pdu(m, i % 2, 900 + i) for i, m in enumerate((1, 3, 6))]
    frames = [O.tx_frame(f, m, 0.2 + 0.1 * i, pd[i], cfo_hz=3.0 * i - 4, phase0=i, amplitude=0.08) for i, (f, m) in enumerate(zip(freqs, (1, 3, 6)))]
    x = O.render(int(sr * 5.9), sr, cf, frames, noise_sigma=O.noise_sigma(0.08, sr, 18.0), seed=5)
    outs = []
    for depth in (0, 1, 3):
        p = O.Pipeline(sr, cf, freqs, fold_mode=O.FOLD_FULL, nthreads=3, ring_depth=depth)
        for c in range(3):
            p.set_capture(c, ["eq"], 1 << 16)
        rng = np.random.default_rng(depth)
        i = 0
        while i < x.size:                                  # ragged feeding: blocks are produced as they fill
            k = int(rng.integers(1, 90000))
            p.feed(x[i:i + k])
            i += k
        got = p.pdus()                                     # waits until the workers have drained the ring
        outs.append(([(q.freq, q.sample_cnt_end, q.data(), q.crc_good) for q in got], [p.capture(c, "eq") for c in range(3)], [p.stats(c) for c in range(3)]))
        p.close()
    assert outs[0][0] == outs[1][0] == outs[2][0] and len(outs[0][0]) == 3
    assert sorted(d for _, _, d, _ in outs[0][0]) == sorted(pd)
    for a, b in ((0, 1), (0, 2)):
        assert all(np.array_equal(u, v) for u, v in zip(outs[a][1], outs[b][1]))
        assert outs[a][2] == outs[b][2]
