"""GPU parity tests (-m gpu): libhfdl_b200.so on a real B200 through the C ABI vs the CPU oracle."""
import json
import os

import numpy as np
import pytest

import b200_cases as K
import golden_cases as GC
import dumphfdl_b200 as hb
import dumphfdl_b200.api as A
import orclib as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    L = hb.load()
    assert L.hfdl_b200_device_count() >= 1
    return L


def test_fft_all_plans(lib):
    # 1-, 2- and 3-pass plans incl. the BASELINE sizes 2^15 (cfg1), 2^18 (cfg2), 2^22 (cfg3/4), 2^23 (cfg5)
    K.case_fft(lib, [256, 4096, 8192, 32768, 262144, 524288], batch=3)
    K.case_fft(lib, [1 << 22], batch=2, seed=2)
    K.case_fft(lib, [1 << 23], batch=1, seed=3)


def test_viterbi_bitexact_all_sizes(lib):
    K.case_viterbi(lib, [540, 1080, 1260, 2160, 2520, 3240, 5040, 7560], frames=5)


def test_fec_bitexact_all_modes(lib):
    K.case_fec(lib, range(8))


def test_fec_roundtrip_at_scale(lib):
    # encode -> decode of 512 frames in one launch: every PDU comes back, FCS good (size-independent property)
    L = O.lib()
    for M1 in (1, 7):
        nsym = [2160, 5040][M1 // 4]
        pd = [O.make_pdu(M1, k % 2, seed=7000 + k) for k in range(512)]
        S = np.zeros((512, nsym), np.complex64)
        sym = np.zeros(5040, np.complex64)
        for k, p in enumerate(pd):
            L.orc_encode_user_data(np.frombuffer(p, np.uint8).copy(), M1, sym)
            S[k] = sym[:nsym]
        out, crc, _ = A.fec_decode(S, M1, 0, lib=lib)
        assert all(bytes(out[k]) == pd[k] for k in range(512)) and crc.all()


def test_frontend_cfg1_cs16(lib):
    assert K.case_frontend(lib, 250000, [10063000], [1], 3.2, sfmt=A.SFMT_CS16, batch=8) == 1


def test_frontend_cu8(lib):
    assert K.case_frontend(lib, 250000, [10063000], [2], 3.2, sfmt=A.SFMT_CU8, batch=5, check_floats=False) == 1


def test_frontend_cfg2_8ch_all_modes(lib):
    # BASELINE config 2: 2 Msps CF32, 8 channels (one frame of every M1 mode)
    sr = 2000000
    freqs = [K.CF + int((k - 3.5) * 212000) // 1000 * 1000 for k in range(8)]
    assert K.case_frontend(lib, sr, freqs, list(range(8)), 5.8, batch=16, seed=11) == 8


def test_frontend_three_pass_fft_plan(lib):
    # 8 Msps: N = 2^20 -> the three-pass forward FFT plan (as at BASELINE configs 3-5) under the whole demodulator
    sr = 8000000
    freqs = [K.CF - 3100000, K.CF + 40000, K.CF + 2900000]
    assert K.case_frontend(lib, sr, freqs, [1, 3, 0], 2.9, batch=5, seed=17) == 3


def test_frontend_cfg3_geometry_many_channels(lib):
    """BASELINE config 3 geometry: 20 Msps -> N = 2^22 (three-pass FFT, 2 blocks per L2-sized sub-batch), M = 4096,
    1792 outputs per block, resampler 0.55296; 40 channels spread over the band, all four single-slot modes.
    PDUs, counters, front parser, spectrum / channeliser / AGC / MF / EQ checkpoints vs the oracle.  The channeliser
    tolerance is wider than at 896 outputs per block: the oracle follows the reference's recursive phasor
    (libcsdr_gpl.c:41-74, whose error grows along the 1792-sample block), the device uses the closed-form phase."""
    sr = 20000000
    nch = 40
    delta = int(0.85 * sr / nch / 1000) * 1000
    freqs = [int(round((K.CF + (k - (nch - 1) / 2) * delta) / 1000.0)) * 1000 for k in range(nch)]
    fe = A.Frontend(sr, K.CF, freqs[:1], max_blocks_per_batch=1, lib=lib)
    g = fe.geom
    assert (g.fft_size, g.fft_inv_size, g.input_size, g.out_per_block, g.fft_passes) == (1 << 22, 4096, 3670016, 1792, 3)
    assert abs(g.resamp_rate - 0.55296) < 1e-6
    fe.close()
    assert K.case_frontend(lib, sr, freqs, [k % 4 for k in range(nch)], 2.8, batch=5, seed=23, starts=[0.05 + 0.005 * k for k in range(nch)], tol_ddc=2e-4) == nch


def test_frontend_cfg4_cfg5_geometry(lib):
    """BASELINE config 4 / 5 geometries: 30 Msps -> N = 2^22, M = 2048 (pre-decimation 2048); 60 Msps CS16 -> N = 2^23
    (plan 128 x 256 x 256), M = 2048, overlap 2^20.  Four channels across the band; PDUs, counters, front parser and the
    float checkpoints vs the oracle."""
    for sr, sfmt, N, isz in ((30000000, A.SFMT_CF32, 1 << 22, 3670016), (60000000, A.SFMT_CS16, 1 << 23, 7340032)):
        freqs = [K.CF + off for off in (-9100000, -2003000, 3907000, 9702000)]
        fe = A.Frontend(sr, K.CF, freqs[:1], sample_format=sfmt, max_blocks_per_batch=1, lib=lib)
        g = fe.geom
        assert (g.fft_size, g.fft_inv_size, g.input_size, g.out_per_block, g.fft_passes) == (N, 2048, isz, 896, 3)
        fe.close()
        # equaliser checkpoint: at N = 2^23 (a million taps) the channeliser outputs of oracle and device differ by ~1e-5
        # instead of ~1e-6 and the decision-directed loops amplify that; AGC and matched filter keep the usual bound
        assert K.case_frontend(lib, sr, freqs, [1, 3, 0, 2], 2.7, sfmt=sfmt, batch=6, seed=29, starts=[0.03 + 0.01 * k for k in range(4)],
                               tol_demod=5e-3 if sr == 60000000 else K.TOL_DEMOD) == 4
    _dump_measured()


def test_loop_kernel_layouts_agree(lib, monkeypatch):
    # loop_kernel's CTA layouts (four channels per CTA; two, role-major, padded shared memory as in multi-GPU mode; two, plain):
    # same PDUs, counters and checkpoints as the oracle for each
    for lay, kb in (("role2", "216"), ("pair2", "0"), ("pack4", "216")):
        monkeypatch.setenv("HFDL_B200_LOOP_LAYOUT", lay)
        monkeypatch.setenv("HFDL_B200_LOOP_SMEM_KB", kb)
        assert K.case_frontend(lib, 2000000, [K.CF + 212000, K.CF - 424000, K.CF + 636000, K.CF - 100000, K.CF + 800000], [3, 1, 0, 2, 1], 2.9, batch=5, seed=19) == 5


def _dump_measured():
    d = os.path.join(os.path.dirname(HERE), "gpurun_out")
    if os.path.isdir(d) and K.MEASURED:
        with open(os.path.join(d, "parity_errors.json"), "w") as f:
            json.dump(K.MEASURED, f, indent=1)


def test_tapslice_checkpoint_vs_oracle(lib):
    # cfg 2 and cfg 3 geometries: the tap spectrum the slice fold multiplies with (fastddc.c:217-252)
    K.case_tapslice(lib, 2000000, [K.CF + 212000, K.CF - 777000])
    K.case_tapslice(lib, 20000000, [K.CF + 8123000, K.CF - 40000, K.CF - 9001000])


def test_pruned_spectrum_equals_full(lib):
    # production mode (only the channels' spectrum granules are stored by the last FFT pass) vs parity mode, cfg 2 and 8 Msps plans
    assert K.case_pruned_spectrum(lib, 2000000, [K.CF + 212000, K.CF - 424000, K.CF + 636000, K.CF - 900000], [3, 1, 0, 2], 2.9, batch=7) == 4
    assert K.case_pruned_spectrum(lib, 8000000, [K.CF - 3100000, K.CF + 40000, K.CF + 3900000], [1, 3, 0], 2.9, batch=5, seed=62) == 3


def test_errors_and_empty_inputs(lib):
    K.case_errors_and_empty_inputs(lib)


def test_front_parser_all_branches(lib):
    K.case_front_parser(lib)


def test_streaming_submit_poll(lib):
    # the block shim's use of the C ABI: hfdl_b200_push_samples + hfdl_b200_submit + hfdl_b200_poll, no flush before the end
    plan = [(0, 0, 0.2), (0, 2, 3.2), (1, 3, 0.5), (1, 1, 3.5), (2, 1, 1.4)]
    assert K.case_frontend_stream(lib, 250000, [10063000, 9952000, 10101000], plan, 6.4, batch=4, push_blocks=1, seed=37, submit_poll=True) >= 4


def test_entry_points_from_other_threads_and_devices(lib):
    """The current CUDA device is per host thread: a frontend created here must work when it is pushed, polled and
    queried from other threads (block.c starts the thread routine on a fresh pthread; a stats thread reads beside it).
    With two GPUs the frontend lives on device 1 while every calling thread's current device is 0."""
    import threading
    dev = 1 if lib.hfdl_b200_device_count() >= 2 else 0
    sr, freqs = 250000, [10063000, 9952000]
    x, truth = K.make_capture(sr, freqs, [1, 2], 3.3, seed=21)
    p = K.run_oracle(sr, freqs, x, O.SFMT_CF32)
    fe = A.Frontend(sr, K.CF, freqs, max_blocks_per_batch=4, device=dev, lib=lib)
    isz = fe.geom.input_size
    err, stop = [], [False]

    def pusher():
        try:
            for i in range(0, x.size, isz):
                fe.push(x[i:i + isz])
                fe.submit()
                fe.poll()
            fe.flush()
        except Exception as e:       # noqa: BLE001
            err.append(e)

    def reader():
        try:
            while not stop[0]:
                for c in range(len(freqs)):
                    assert fe.noise_floor(c) > 0
                    fe.counters(c)
        except Exception as e:       # noqa: BLE001
            err.append(e)

    t1, t2 = threading.Thread(target=pusher), threading.Thread(target=reader)
    t2.start()
    t1.start()
    t1.join()
    stop[0] = True
    t2.join()
    assert not err, err
    K.compare_pdus(fe.pdus(), p.pdus(), truth)
    K.check_counters(fe, p, freqs, p.pdus())
    fe.close()


def test_two_gpus_peer_broadcast_equals_single(lib):
    """Channels k mod 2 on two GPUs, the capture pushed to GPU 0 only and copied ring to ring over NVLink
    (hfdl_b200_push_peer): the union of the PDUs equals the oracle's list."""
    if lib.hfdl_b200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sr = 2000000
    freqs = [K.CF + 212000, K.CF - 424000, K.CF + 636000, K.CF - 100000]
    x, truth = K.make_capture(sr, freqs, [3, 5, 0, 2], 5.8, seed=35)
    ref = K.run_oracle(sr, freqs, x, O.SFMT_CF32).pdus()
    fes = [A.Frontend(sr, K.CF, freqs[d::2], max_blocks_per_batch=6, device=d, lib=lib) for d in range(2)]
    isz = fes[0].geom.input_size
    got = []
    for i in range(0, x.size, 3 * isz):
        fes[0].push(x[i:i + 3 * isz])
        fes[1].push_peer(fes[0])
        for f in fes:
            f.submit()
            f.poll()
            got += f.pdus()
    for f in fes:
        f.flush()
        got += f.pdus()
        f.close()
    K.compare_pdus(got, ref, truth)


def test_frontend_low_snr_equals_oracle(lib):
    """Es/N0 = 8 and 5 dB: decisions are marginal and some payloads come back with bit errors behind a good header FCS;
    the GPU path must still produce exactly the oracle's PDU list, counters and float checkpoints."""
    for esn0 in (8.0, 5.0):
        n = K.case_frontend(lib, 250000, [10063000, 9952000, 10101000, 9900000], [0, 1, 2, 3], 3.4, batch=6, seed=41, esn0=esn0, check_truth=False)
        assert n >= 3


def test_frontend_ragged_push_and_batch_size_invariance(lib):
    n1 = K.case_frontend(lib, 250000, [10063000, 9952000, 10101000], [3, 0, 5], 5.6, batch=3, ragged=True, seed=5)
    n2 = K.case_frontend(lib, 250000, [10063000, 9952000, 10101000], [3, 0, 5], 5.6, batch=64, seed=5)
    assert n1 == n2 == 3


def test_ring_protocol_under_small_launches_repeated(lib):
    """Many short loop_kernel launches (the consumer warp runs right behind the producer warp after every launch
    start): the continuous equaliser checkpoint must stay within tolerance on every repetition.  This is the case
    that exposed torn 16-byte ring entries."""
    for rs in (1, 2, 3, 4):
        assert K.case_frontend(lib, 250000, [10063000, 9952000, 10101000], [3, 0, 5], 5.6, batch=3, ragged=True, seed=5, ragged_seed=rs) == 3


def test_pipeline_many_batches_subranges(lib):
    # 5 batches of 32 blocks (n_out > 16384: the 8 sub-range schedule) with frames in every batch, whole capture in one push
    plan = [(0, 1, 0.3), (0, 3, 3.3), (0, 0, 6.3), (0, 2, 9.3), (0, 1, 12.3), (0, 3, 15.3),
            (1, 2, 0.9), (1, 5, 3.9), (1, 0, 9.9), (1, 7, 12.9)]
    assert K.case_frontend_stream(lib, 250000, [10063000, 9952000], plan, 18.4, batch=32) >= 8


def test_pipeline_small_batches_streaming_pickup(lib):
    # 2-block batches pushed 3 blocks at a time, PDUs popped between pushes: deferred collection keeps order and count
    plan = [(0, 0, 0.2), (0, 2, 3.2), (1, 3, 0.5), (1, 1, 3.5), (2, 1, 1.4)]
    assert K.case_frontend_stream(lib, 250000, [10063000, 9952000, 10101000], plan, 6.4, batch=2, push_blocks=3, seed=33) >= 4


def test_device_resident_path_equals_host_path(lib):
    import torch
    sr, freqs = 250000, [10063000, 9952000]
    x, truth = K.make_capture(sr, freqs, [1, 2], 3.3, seed=21)
    fe1 = A.Frontend(sr, K.CF, freqs, max_blocks_per_batch=8, lib=lib)
    fe1.push(x)
    fe1.flush()
    a = fe1.pdus()
    fe2 = A.Frontend(sr, K.CF, freqs, max_blocks_per_batch=8, lib=lib)
    isz = fe2.geom.input_size
    nb = x.size // isz
    d = torch.from_numpy(x.view(np.float32).copy()).cuda()
    torch.cuda.synchronize()
    done = 0
    while done < nb:
        k = min(5, nb - done)
        fe2.process_device(d.data_ptr(), x.size, done * isz, k)
        done += k
    fe2.sync()                  # process_device only queues the batches
    b = fe2.pdus()
    K.compare_pdus(b, [O_pdu(q) for q in a], truth)


def test_sharded_spectrum_equals_single_frontend(lib):
    # multi-GPU data path on one device: 2 and 4 "ranks", CF32 and CS16, batches that end ragged
    assert K.case_sharded_spectrum(lib, K.CudaMem(), 250000, [10063000, 9952000, 10101000, 9931000], [1, 2, 0, 3], 3.3, nranks=2, batch=4) == 4
    assert K.case_sharded_spectrum(lib, K.CudaMem(), 250000, [10063000, 9952000, 10101000, 9931000], [5, 2, 4, 3], 5.6, nranks=4, batch=8, sfmt=A.SFMT_CS16, seed=72) == 4
    # the exchange fused into the pack kernel (hfdl_b200_spectrum_slices_to)
    assert K.case_sharded_spectrum(lib, K.CudaMem(), 250000, [10063000, 9952000, 10101000, 9931000], [5, 2, 4, 3], 5.6, nranks=4, batch=8, sfmt=A.SFMT_CS16, seed=72, direct=True) == 4


def test_sharded_spectrum_cfg3_geometry(lib):
    # three-pass natural-order plan (N = 2^22, M = 4096) with the pruned spectrum store, 2 ranks x 4 channels
    sr = 20000000
    freqs = [K.CF + int((k - 3.5) * 132000) // 1000 * 1000 for k in range(8)]
    assert K.case_sharded_spectrum(lib, K.CudaMem(), sr, freqs, [1, 2, 0, 3, 1, 0, 2, 3], 2.8, nranks=2, batch=4, seed=73, starts=[0.05 + 0.005 * k for k in range(8)]) == 8


class O_pdu:
    def __init__(self, q):
        self.freq, self.sample_cnt_end, self._d, self.M1, self.crc_good, self.sample_cnt_a2 = q.freq, q.sample_cnt_end, q.data(), q.M1, q.crc_good, q.sample_cnt_a2

    def data(self):
        return self._d


@pytest.mark.parametrize("name", GC.FIXTURES)
def test_golden_fixture(lib, name):
    """PDU octets, hfdl_pdu_metadata fields and statsd counters recorded from the reference's own code in the development
    container (tests/golden/make_golden.py), frame positions from the oracle."""
    G = GC.load(name)
    pdus, counters = GC.run_frontend(G, lib)
    GC.check_against(G, pdus, counters)


def test_no_cpu_fallback_symbols(lib):
    # the product library must not export anything of the oracle
    import subprocess
    out = subprocess.run(["nm", "-D", hb.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out and "hfdl_b200_create" in out


# Added after the GPU time of round 2 was spent: green under host emulation (test_cusim_logic.py), not yet run on a GPU, so
# they do not gate the suite until they have been seen there once (HFDL_B200_EXTRA=1 pytest tests -m gpu).  DESIGN.md 4
# describes the one kind of event (a timing-arm / slicer decision on a float boundary) that may need a tolerance here.
extra = pytest.mark.skipif(not os.environ.get("HFDL_B200_EXTRA"), reason="opt-in: HFDL_B200_EXTRA=1 (not yet run on a GPU)")


@extra
@pytest.mark.parametrize("name", K.HOSTILE)
def test_hostile_captures_follow_the_oracle(lib, name):
    K.case_hostile(lib, name)


@extra
@pytest.mark.parametrize("seed", [0, 2, 5, 11, 15])
def test_random_jobs_follow_the_oracle(lib, seed):
    assert K.case_random_job(lib, seed) >= 2

