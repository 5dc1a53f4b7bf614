"""Generates tests/golden/*.json in the development container (where /root/reference exists).

Known answers for deterministic transmitted captures (BASELINE config 1 geometry: 250 ksps CS16), recorded from THE
REFERENCE'S OWN CODE run here -- block.c + fft.c + fastddc.c (all-bin fold) + hfdl.c + viterbi27_port.c wired as main.c
wires them (oracle/_ref/libref.so; the liquid-dsp objects it calls are served by the oracle), then its own
pdu_decoder_thread / mpdu.c / spdu.c / lpdu.c for the frame counters (oracle/_ref/libref_front.so):
  pdus[].octets, pdus[].meta   PDU octets and the hfdl_pdu_metadata fields (hfdl.c:1061-1067), per-channel order
  statsd                       the per-channel counters the reference fires (hfdl.c:818,828,840; pdu.c:123; mpdu.c / spdu.c / lpdu.c)
plus frame positions (A2 sample, end sample), M1 and the header-FCS flag from the oracle, which the reference does not
export.  The oracle (both folds) must agree with the reference before anything is written.
  cfg1_pdus.json    two channels, one frame each (modes 1 and 6), Es/N0 20 dB
  cfg1_multi.json   four channels, all eight modes (a single-slot and a double-slot frame per channel), MPDUs with good /
                    bad / short / truncated LPDUs, SPDUs and random octets, Es/N0 12 dB
Run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import golden_cases as GC  # noqa: E402
import orclib as O  # noqa: E402

DEMOD = ("demod.preamble.A2_found", "demod.preamble.M1_found", "demod.preamble.errors.M1_not_found")


def record(G, name, expect_all=True):
    raw = GC.capture_of(G, verify=False)
    G["int16_checksum"] = GC.checksum(raw)
    # ---- the reference itself
    r = O.RefPipeline(G["sample_rate"], GC.CF, G["freqs"], sfmt=O.SFMT_CS16, fft_threads=2)
    r.feed(raw)
    r.finish()
    ref = r.pdus()
    statsd = {str(f): {n: r.stat(f, n) for n in DEMOD} for f in G["freqs"]}
    r.close()
    for f in G["freqs"]:
        for c in O.ref_front_run([q.data() for q in ref if q.freq == f], freq=f):
            for n, v in c.items():
                if not n.startswith("nodes.") and n != "other":
                    statsd[str(f)][n] = statsd[str(f)].get(n, 0) + v
    # ---- the oracle, for the fields the reference keeps to itself; it must agree on everything else
    out = {}
    for fold, mode in (("slice", O.FOLD_SLICE), ("full", O.FOLD_FULL)):
        p = O.Pipeline(G["sample_rate"], GC.CF, G["freqs"], fold_mode=mode, nthreads=4)
        p.feed(raw, O.SFMT_CS16)
        out[fold] = [dict(freq=q.freq, a2=int(q.sample_cnt_a2), end=int(q.sample_cnt_end), M1=q.M1, crc_good=q.crc_good,
                          octets=q.data().hex()) for q in p.pdus()]
    assert out["slice"] == out["full"], "slice fold and the reference's full fold must give the same PDUs"
    pdus = out["slice"]
    for f in G["freqs"]:        # per channel the reference's order is fixed (across channels it is thread-race order)
        a = [q for q in ref if q.freq == f]
        b = [d for d in pdus if d["freq"] == f]
        assert len(a) == len(b), (f, len(a), len(b))
        for q, d in zip(a, b):
            assert q.data().hex() == d["octets"], "the reference's PDU octets differ from the oracle's"
            d["meta"] = dict(version=q.version, bit_rate=q.bit_rate, slot=q.slot.decode(), freq_err_hz=float(q.freq_err_hz),
                             rssi=float(q.rssi), noise_floor=float(q.noise_floor))
    G["pdus"] = pdus
    G["statsd"] = statsd
    G["recorded_from"] = ("the reference's own sources run in the development container (oracle/_ref/libref.so, libref_front.so); "
                          "a2 / end / M1 / crc_good from the oracle")
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(G, f, indent=1)
    print(name, "wrote", len(pdus), "PDUs,", sum(d["crc_good"] for d in pdus), "with a good header FCS")
    return pdus


# ---- fixture 1: the capture of b200_cases.make_capture (kept from round 1)
G1 = dict(sample_rate=250000, freqs=[10063000, 9952000], modes=[1, 6], dur=5.6, seed=77)
p1 = record(G1, "cfg1_pdus.json")
assert len(p1) == 2 and all(d["crc_good"] for d in p1)

# ---- fixture 2: every mode, every PDU kind of the transmitter (orc_tx_make_pdu: 0 downlink MPDU, 1 SPDU, 2 random octets,
# 3 uplink MPDU with a bad LPDU, 4 downlink MPDU with short / truncated LPDUs)
rng = np.random.default_rng(2024)
freqs = [10063000, 9952000, 10101000, 9931000]
plan = []
for ch in range(4):
    cfo, ph = float(np.round(rng.uniform(-12, 12), 3)), float(np.round(rng.uniform(0, 6.28), 3))
    plan.append(dict(ch=ch, M1=ch, start=0.2 + 0.05 * ch, kind=[0, 1, 3, 4][ch], pdu_seed=9100 + ch, cfo_hz=cfo, phase0=ph))
    plan.append(dict(ch=ch, M1=4 + ch, start=3.9 + 0.05 * ch, kind=[4, 3, 2, 0][ch], pdu_seed=9200 + ch, cfo_hz=cfo, phase0=ph))
G2 = dict(sample_rate=250000, freqs=freqs, plan=plan, dur=9.6, seed=78, esn0_db=12.0, amplitude=0.1)
p2 = record(G2, "cfg1_multi.json")
assert len(p2) == 8, "every transmitted frame is expected back"
