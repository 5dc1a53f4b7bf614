"""CPU logic tests of the CUDA kernels: dumphfdl_b200/csrc compiled for host emulation (tests/cusim) and run
through the same C ABI and the same parity cases as the GPU tests, at small sizes.  This is test
infrastructure (there is no GPU in the development container); the product never loads this library."""
import ctypes as C
import os
import subprocess

import pytest

import b200_cases as K
import dumphfdl_b200.api as A

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim")], check=True)
    return A.bind(C.CDLL(os.path.join(HERE, "cusim", "libhfdl_cusim.so")))


def test_fft_one_and_two_pass(sim):
    K.case_fft(sim, [64, 2048, 8192, 32768])


def test_fft_every_power_of_two_plan(sim):
    # every FFT size fastddc_init can ask for (2^15 at 250 ksps ... 2^23 at 60 Msps) and the smaller ones the stage entry
    # accepts: one-, two- and three-pass plans with every pass length make_plan produces (frontend.cu)
    K.case_fft(sim, [1 << k for k in range(6, 19)], batch=2, seed=11)
    K.case_fft(sim, [1 << k for k in range(19, 24)], batch=1, seed=12)


def test_viterbi_bitexact(sim):
    K.case_viterbi(sim, [540, 1080], frames=2)


def test_fec_bitexact(sim):
    K.case_fec(sim, [0, 3, 6])


def test_frontend_cfg1_cs16(sim):
    # BASELINE config 1: 250 ksps CS16, one channel
    assert K.case_frontend(sim, 250000, [10063000], [1], 3.2, sfmt=A.SFMT_CS16, batch=4) == 1


def test_frontend_other_sample_rates(sim):
    # geometries between the BASELINE configurations: 1.024 Msps (N = 2^18 with M = 4096, resampler 0.675) and 768 ksps
    # (N = 2^17, M = 2048, resampler 0.9), channels near the band edge
    assert K.case_frontend(sim, 1024000, [K.CF + 401000], [2], 3.2, sfmt=A.SFMT_CS16, batch=3, seed=12, tol_ddc=2e-4) == 1
    assert K.case_frontend(sim, 768000, [K.CF - 333000], [1], 3.2, batch=3, seed=13) == 1


def test_frontend_two_channels_ragged_cf32(sim):
    assert K.case_frontend(sim, 250000, [10063000, 9952000], [3, 0], 3.2, batch=3, ragged=True, seed=5) == 2


def test_frontend_large_batch_uses_subranges(sim):
    # one 48-block batch: n_out > 16384 -> the agc/bank || loop sub-range schedule (HFDL_NSUB) is exercised
    assert K.case_frontend(sim, 250000, [10063000, 9952000], [5, 1], 5.6, batch=64, seed=8) == 2


def test_stream_small_batches_pickup(sim):
    # host-emulation run of the streaming case (pipeline bookkeeping: set alternation, deferred collection, frame slots)
    plan = [(0, 0, 0.2), (0, 1, 3.2)]
    assert K.case_frontend_stream(sim, 250000, [10063000], plan, 6.2, batch=2, push_blocks=3, seed=33) == 2


def test_stream_submit_poll(sim):
    # the block shim's use of the C ABI: push + hfdl_b200_submit + hfdl_b200_poll, never a flush before the end
    plan = [(0, 1, 0.2), (1, 2, 0.6)]
    assert K.case_frontend_stream(sim, 250000, [10063000, 9952000], plan, 3.4, batch=4, push_blocks=1, seed=35, submit_poll=True) == 2


def test_errors_and_empty_inputs(sim):
    K.case_errors_and_empty_inputs(sim)


@pytest.mark.parametrize("name", K.HOSTILE)
def test_hostile_captures_follow_the_oracle(sim, name):
    # collisions (one gives a PDU with a bad FCS), carrier offsets up to no detection at all, clipping, frames cut by the end /
    # the start of the capture, a strong adjacent carrier, a DC spur with a weak frame
    # (the oracle equals the reference's own hfdl.c on these captures: test_oracle_hfdl_ref.py)
    K.case_hostile(sim, name)


@pytest.mark.parametrize("seed", [11, 15])
def test_random_jobs_follow_the_oracle(sim, seed):
    # two of the seeded random jobs of test_oracle_hfdl_ref.py::test_reference_hfdl_equals_oracle_on_random_jobs
    assert K.case_random_job(sim, seed) >= 2


def test_front_parser(sim):
    K.case_front_parser(sim)


def test_tapslice_checkpoint(sim):
    K.case_tapslice(sim, 250000, [10063000, 9931000])


def test_pruned_spectrum_equals_full(sim):
    assert K.case_pruned_spectrum(sim, 250000, [10063000, 9931000, 10110000], [1, 2, 0], 3.3, batch=5) == 3


def test_sharded_spectrum_two_ranks(sim):
    # multi-GPU data path (FFT blocks sharded, slices exchanged, channels sharded) on the host emulation: PDUs == one frontend's.
    # Here with the exchange fused into the pack kernel (stores straight into every rank's receive buffer); the packed send
    # buffer + all-to-all variant runs in test_multi_rank.py (gloo, two processes) and test_block_shim.py
    assert K.case_sharded_spectrum(sim, K.HostMem(), 250000, [10063000, 9952000, 10101000, 9931000], [1, 2, 0, 3], 3.3, nranks=2, batch=4, direct=True) == 4


def test_loop_kernel_layouts_agree(sim, monkeypatch):
    # the three CTA layouts of loop_kernel (pack4: four channels per CTA; role2: two, role-major with idle warps; pair2) give the same PDUs
    for lay in ("role2", "pack4"):
        monkeypatch.setenv("HFDL_B200_LOOP_LAYOUT", lay)
        assert K.case_frontend(sim, 250000, [10063000, 9952000], [3, 0], 3.2, batch=3, seed=6) == 2


def test_push_nowait_two_staging_buffers(sim):
    # a producer with two staging buffers: hfdl_b200_push_samples_nowait + hfdl_b200_wait_host_buffer before a buffer is refilled
    import numpy as np
    sr, freqs = 250000, [10063000, 9952000]
    x, truth = K.make_capture(sr, freqs, [1, 2], 3.3, seed=27)
    ref = K.run_oracle(sr, freqs, x, A.SFMT_CF32).pdus()
    fe = A.Frontend(sr, K.CF, freqs, max_blocks_per_batch=4, lib=sim)
    chunk = 3 * fe.geom.input_size + 1234
    stage = [np.zeros(chunk, np.complex64), np.zeros(chunk, np.complex64)]
    got = []
    for i, at in enumerate(range(0, x.size, chunk)):
        seg = x[at:at + chunk]
        buf = stage[i % 2]
        if i >= 2:
            fe.wait_host_buffer()                 # (one event for the newest copy: waiting for it covers the older ones)
        buf[:seg.size] = seg
        fe.push_ptr(buf.ctypes.data, seg.size, wait=False)
        got += fe.pdus()
    fe.wait_host_buffer()
    fe.flush()
    got += fe.pdus()
    K.compare_pdus(got, ref, truth)
    fe.close()
