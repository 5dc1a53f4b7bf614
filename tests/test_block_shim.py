"""The block.c-facing wrapper (include/hfdl_b200_block.h) against the REFERENCE'S OWN block.c: tests/block_driver.c loads
oracle/_ref/libref.so (the reference's block.c + input-helpers.c compiled where they lie, a cbuffercf with liquid-dsp's
API, pdu_decoder_queue_push as a capture list) and then the front-end library, wires input -> front-end with
block_connect_one2one / block_start exactly as main.c does, replays a capture file and prints what
pdu_decoder_queue_push received.  CPU variant: host-emulation build; GPU variants: the real library, 1 and 2 GPUs."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import b200_cases as K
import orclib as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBREF = os.path.join(ROOT, "oracle", "_ref", "libref.so")
BITRATE = {0: 300, 1: 600, 2: 1200, 3: 1800}

pytestmark = pytest.mark.skipif(not os.path.exists(LIBREF) and not os.path.isdir("/root/reference/src"),
                                reason="oracle/_ref/libref.so (reference block.c compiled in place) not available")


def run_driver(libpath, sr, freqs, modes, dur, seed, ngpus=1, env=None):
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim"), "all"], check=True)
    O.reflib()
    x, truth = K.make_capture(sr, freqs, modes, dur, seed=seed)
    p = K.run_oracle(sr, freqs, x, O.SFMT_CF32)
    ref = p.pdus()
    with tempfile.NamedTemporaryFile(suffix=".cf32", delete=False) as f:
        x.tofile(f)
        path = f.name
    try:
        out = subprocess.run([os.path.join(HERE, "block_driver"), LIBREF, libpath, path, str(sr), str(K.CF), str(ngpus)] + [str(f) for f in freqs],
                             capture_output=True, text=True, timeout=900, check=True, env=dict(os.environ, **(env or {}))).stdout
    finally:
        os.unlink(path)
    rows = [l.split() for l in out.splitlines()]
    got = sorted((int(a[1]), int(a[2]), a[3], a[8]) for a in rows if a and a[0] == "PDU")
    want = sorted((q.freq, BITRATE[q.M1 % 4], "S" if q.M1 < 4 else "D", q.data().hex()) for q in ref)
    assert got == want                                   # what pdu_decoder_queue_push received == the oracle's PDUs
    assert sorted((g[0], bytes.fromhex(g[3])) for g in got) == sorted(truth)
    # metadata as struct hfdl_pdu_metadata carries it (hfdl.c:1061-1067), read back through the reference's own pdu.h layout
    byfreq = {q.freq: q for q in ref}
    for a in rows:
        if a and a[0] == "PDU":
            q = byfreq[int(a[1])]
            assert int(a[4]) == 1
            assert abs(float(a[5]) - q.freq_err_hz) < 1e-2
            assert abs(float(a[6]) - 20 * np.log10(q.signal_level)) < 1e-2
            assert abs(float(a[7]) - 20 * np.log10(q.noise_floor)) < 1e-2
    # the statsd-style counters read while the blocks ran and at the end (doc/STATSD_METRICS.md)
    cnt = {int(a[1]): [int(v) for v in a[2:]] for a in rows if a and a[0] == "CNT"}
    for i, f in enumerate(freqs):
        a1, a2, m1, frames = p.stats(i)
        mine = [q for q in ref if q.freq == f]
        fr = [O.pdu_front(q.data()) for q in mine]
        assert cnt[f][0:3] == [a2, m1, p.m1_not_found(i)]
        assert cnt[f][3] == len(mine) and cnt[f][4] == sum(1 for v in fr if v[0] == 0) and cnt[f][5] == sum(1 for v in fr if v[0] == 1)
        assert cnt[f][6] == sum(1 for v in fr if v[0] == 0 and v[1] == 1) and cnt[f][7] == sum(1 for v in fr if v[0] == 0 and v[1] == 0)
        assert cnt[f][8] == sum(v[2] for v in fr) and cnt[f][9] == sum(v[3] for v in fr)
    # ... and the demod.preamble.* increments the block forwarded to the host program's statsd hook (hfdl.c:818,828,840)
    sd = {int(a[1]): [int(v) for v in a[2:]] for a in rows if a and a[0] == "STATSD"}
    for f in freqs:
        assert sd[f] == cnt[f][0:3], (f, sd[f], cnt[f][0:3])
    assert int([a for a in rows if a and a[0] == "POLLS"][0][1]) > 0
    return len(got)


def test_block_wrapper_error_behaviour_host_emulation():
    """hfdl_gpu_frontend_create refuses what main.c refuses before it builds its blocks (no frequency, a channel outside the
    capture's span, main.c:214-226,687-695) and devices that are not there; the accessors take a NULL block and bad indices"""
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim"), "all"], check=True)
    O.reflib()
    out = subprocess.run([os.path.join(HERE, "block_driver"), LIBREF, os.path.join(HERE, "cusim", "libhfdl_cusim.so"), "badargs",
                          "250000", str(K.CF), "1", "10063000", "9952000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "BADARGS 0" in out.stdout, (out.stdout, out.stderr[-2000:])


def test_block_contract_host_emulation():
    lib = os.path.join(HERE, "cusim", "libhfdl_cusim.so")
    assert run_driver(lib, 250000, [10063000, 9952000], [1, 2], 3.3, seed=31) == 2


def test_block_contract_host_emulation_two_devices_sharded_spectrum():
    # ngpus = 2 on the host emulation (two "devices" sharing the host memory): the sharded-spectrum mode of the wrapper --
    # each GPU transforms its share of every batch's blocks, slices change hands by peer copies, channels are sharded
    lib = os.path.join(HERE, "cusim", "libhfdl_cusim.so")
    assert run_driver(lib, 250000, [10063000, 9952000, 10101000, 9931000], [1, 2, 0, 3], 3.3, seed=37, ngpus=2, env={"HFDL_CUSIM_DEVICES": "2"}) == 4
    # ... with packed send buffers + device-to-device copies instead of the pack kernel storing into the peers' buffers
    assert run_driver(lib, 250000, [10063000, 9952000, 10101000, 9931000], [1, 2, 0, 3], 3.3, seed=37, ngpus=2, env={"HFDL_CUSIM_DEVICES": "2", "HFDL_B200_SHIM_COPY": "1"}) == 4


@pytest.mark.gpu
def test_block_contract_gpu():
    lib = os.path.join(ROOT, "dumphfdl_b200", "libhfdl_b200.so")
    assert run_driver(lib, 2000000, [K.CF + 212000, K.CF - 424000, K.CF + 636000], [3, 5, 0], 5.8, seed=33) == 3


@pytest.mark.gpu
def test_block_contract_two_gpus():
    import dumphfdl_b200 as hb
    if hb.load().hfdl_b200_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    lib = os.path.join(ROOT, "dumphfdl_b200", "libhfdl_b200.so")
    freqs, modes = [K.CF + 212000, K.CF - 424000, K.CF + 636000, K.CF - 100000], [3, 5, 0, 2]
    # sharded spectrum (4 channels on 2 GPUs: the default), then the capture broadcast (3 channels do not divide by 2; and forced)
    assert run_driver(lib, 2000000, freqs, modes, 5.8, seed=35, ngpus=2) == 4
    assert run_driver(lib, 2000000, freqs[:3], modes[:3], 5.8, seed=36, ngpus=2) == 3
    assert run_driver(lib, 2000000, freqs, modes, 5.8, seed=35, ngpus=2, env={"HFDL_B200_SHIM_BROADCAST": "1"}) == 4
    assert run_driver(lib, 2000000, freqs, modes, 5.8, seed=35, ngpus=2, env={"HFDL_B200_SHIM_COPY": "1"}) == 4
