"""Generates tests/golden/*.json in the development container: known answers recorded from the CPU oracle
(whose channeliser / Viterbi / CRC are pinned against the reference's own sources in oracle/_ref, see
tests/test_oracle_ref.py) for a deterministic transmitted capture.  Run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import b200_cases as K  # noqa: E402
import orclib as O  # noqa: E402

G = dict(sample_rate=250000, freqs=[10063000, 9952000], modes=[1, 6], dur=5.6, seed=77)
x, truth = K.make_capture(G["sample_rate"], G["freqs"], G["modes"], G["dur"], seed=G["seed"])
raw = np.zeros(2 * x.size, np.int16)
O.lib().orc_quantize_cs16(x, x.size, raw)
out = {}
for name, mode in (("slice", O.FOLD_SLICE), ("full", O.FOLD_FULL)):
    p = O.Pipeline(G["sample_rate"], K.CF, G["freqs"], fold_mode=mode, nthreads=4)
    p.feed(raw, O.SFMT_CS16)
    out[name] = [dict(freq=q.freq, a2=int(q.sample_cnt_a2), end=int(q.sample_cnt_end), M1=q.M1, crc_good=q.crc_good,
                      octets=q.data().hex()) for q in p.pdus()]
assert out["slice"] == out["full"], "slice fold and the reference's full fold must give the same PDUs"
assert sorted((d["freq"], bytes.fromhex(d["octets"])) for d in out["slice"]) == sorted(truth)
G["pdus"] = out["slice"]
G["capture_sha_note"] = "capture regenerated deterministically by b200_cases.make_capture(seed)"
G["int16_checksum"] = int(np.bitwise_xor.reduce(raw.view(np.uint16).astype(np.uint64) * np.arange(1, raw.size + 1, dtype=np.uint64) % 65521))
with open(os.path.join(HERE, "cfg1_pdus.json"), "w") as f:
    json.dump(G, f, indent=1)
print("wrote", len(G["pdus"]), "PDUs")
