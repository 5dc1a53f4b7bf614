"""Golden fixtures (tests/golden/*.json): known answers recorded from the reference's own code run in the development
container (tests/golden/make_golden.py).  A fixture names its capture by parameters only (every frame of the plan, the noise
seed): capture_of() regenerates it with the oracle's transmitter, `int16_checksum` guards that regeneration, and
check_against() compares whatever produced PDUs (the oracle, the host-emulation build, the GPU library) with the record."""
import json
import os

import numpy as np

import orclib as O

HERE = os.path.dirname(os.path.abspath(__file__))
CF = 10000000
FIXTURES = ("cfg1_pdus.json", "cfg1_multi.json")


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def checksum(raw):
    return int(np.bitwise_xor.reduce(raw.view(np.uint16).astype(np.uint64) * np.arange(1, raw.size + 1, dtype=np.uint64) % 65521))


def capture_of(G, verify=True):
    """-> CS16 capture (int16 I/Q pairs) of a fixture, regenerated from its plan (verify: against the recorded checksum)"""
    if "plan" not in G:                 # cfg1_pdus.json: the capture of b200_cases.make_capture
        import b200_cases as K
        x, _ = K.make_capture(G["sample_rate"], G["freqs"], G["modes"], G["dur"], seed=G["seed"])
    else:
        frames = [O.tx_frame(G["freqs"][fr["ch"]], fr["M1"], fr["start"], O.make_pdu(fr["M1"], fr["kind"], fr["pdu_seed"]),
                             cfo_hz=fr["cfo_hz"], phase0=fr["phase0"], amplitude=G["amplitude"]) for fr in G["plan"]]
        x = O.render(int(G["sample_rate"] * G["dur"]), G["sample_rate"], CF, frames,
                     noise_sigma=O.noise_sigma(G["amplitude"], G["sample_rate"], G["esn0_db"]), seed=G["seed"])
    raw = np.zeros(2 * x.size, np.int16)
    O.lib().orc_quantize_cs16(x, x.size, raw)
    assert not verify or checksum(raw) == G["int16_checksum"], "the capture of the fixture could not be regenerated"
    return raw


def check_against(G, pdus, counters=None, meta=True):
    """pdus: records with freq / sample_cnt_a2 / sample_cnt_end / M1 / crc_good / data() (+ metadata fields when meta);
    counters: optional {freq: {statsd name: count}} of the implementation under test"""
    got = sorted(pdus, key=lambda q: (q.freq, q.sample_cnt_end))
    want = sorted(G["pdus"], key=lambda d: (d["freq"], d["end"]))
    assert [(q.freq, q.sample_cnt_a2, q.sample_cnt_end, q.M1, q.crc_good, q.data().hex()) for q in got] == \
           [(d["freq"], d["a2"], d["end"], d["M1"], d["crc_good"], d["octets"]) for d in want]
    if meta:
        for q, d in zip(got, want):
            m = d["meta"]                      # struct hfdl_pdu_metadata as the reference's dispatch_pdu filled it (hfdl.c:1061-1067)
            slot = q.slot.decode() if isinstance(q.slot, bytes) else q.slot
            assert (q.bit_rate, slot) == (m["bit_rate"], m["slot"])
            assert abs(q.freq_err_hz - m["freq_err_hz"]) < 1e-2
            assert abs(q.rssi - m["rssi"]) < 1e-2 and abs(q.noise_floor - m["noise_floor"]) < 1e-2          # dB
    if counters is not None:
        for f, want_c in G["statsd"].items():
            for name, v in want_c.items():
                assert counters[int(f)][name] == v, (f, name, counters[int(f)][name], v)


STATSD_OF_COUNTERS = {      # hfdl_b200_counters_t field -> the reference's statsd metric (doc/STATSD_METRICS.md)
    "A2_found": "demod.preamble.A2_found", "M1_found": "demod.preamble.M1_found", "M1_not_found": "demod.preamble.errors.M1_not_found",
    "frames_processed": "frames.processed", "frames_good": "frames.good", "frames_bad_fcs": "frame.errors.bad_fcs",
    "frames_too_short": "frame.errors.too_short", "frames_air2gnd": "frame.dir.air2gnd", "frames_gnd2air": "frame.dir.gnd2air",
    "lpdus_processed": "lpdus.processed", "lpdus_good": "lpdus.good", "lpdus_bad_fcs": "lpdu.errors.bad_fcs",
    "lpdus_too_short": "lpdu.errors.too_short"}


def run_frontend(G, lib, batch=7):
    """the fixture through the C ABI (GPU library or host-emulation build) -> (pdus, {freq: {statsd name: count}})"""
    import dumphfdl_b200.api as A
    raw = capture_of(G)
    fe = A.Frontend(G["sample_rate"], CF, G["freqs"], sample_format=A.SFMT_CS16, max_blocks_per_batch=batch, lib=lib)
    fe.push(raw)
    fe.flush()
    pdus = fe.pdus()
    counters = {}
    for c, f in enumerate(G["freqs"]):
        k = fe.counters(c)
        counters[f] = {name: getattr(k, field) for field, name in STATSD_OF_COUNTERS.items()}
    fe.close()
    return pdus, counters
