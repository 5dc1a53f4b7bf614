"""bench.py's synthetic slab: the cfg4 / cfg5 construction (frames rendered for one sub-band of `group` channels,
the other channels being frequency-shifted copies that stay continuous across the wrap of the cyclic slab) must give
every channel a decodable frame: a miniature with the same construction goes through the oracle, two slab passes."""
import importlib.util
import os

import numpy as np

import orclib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_grouped_slab_every_channel_decodes_across_the_wrap():
    B = _bench()
    W = dict(sr=2000000, nch=12, blocks_per_slot=22, slots=1, seed=4242, loops=1, group=4)
    isz = O.geometry(W["sr"])[2].input_size
    P = B.plan_frames(O, W, isz)
    assert len(P["freqs"]) == 12 and P["ngroups"] == 3 and len(P["truth"]) == 12
    d = np.diff(P["freqs"])
    assert (d == d[0]).all() and d[0] == P["delta"]                  # uniform 1 kHz-grid spacing: copies land on channel centres
    n = P["nsamp"]
    half = n // 2
    x = np.concatenate([B.render_range(O, W, P, 0, half, 4, 1), B.render_range(O, W, P, half, n - half, 4, 2)])
    p = O.Pipeline(W["sr"], B.CF, P["freqs"], fold_mode=O.FOLD_SLICE, nthreads=8)
    p.feed(x)
    p.feed(x)                                                        # second pass: frames that wrap around the slab end complete here
    got = {(q.freq, q.data()) for q in p.pdus() if q.crc_good}
    assert got == set(P["truth"])


def test_config_dict_is_the_same_for_both_arms():
    B = _bench()
    W = B.WORKLOADS["cfg3"]
    isz = O.geometry(W["sr"])[2].input_size
    P = dict(nblocks=14, nsamp=14 * isz)
    assert B.config_of("cfg3", W, P, 1) == B.config_of("cfg3", W, P, 1)
    assert "cfg3" in B.config_of("cfg3", W, P, 1)["workload"]
    assert B.DEFAULT_BY_GPUS == {1: "cfg3", 2: "cfg3", 4: "cfg4", 8: "cfg5"}


def test_workload_per_gpu_count():
    """prepare_workload: BASELINE's configuration for the GPU count, CS16 for the wideband ones, a slab whose block count
    divides by the GPU count, the sharded spectrum unless the channels do not divide; both arms call it the same way."""
    import argparse
    B = _bench()
    a = argparse.Namespace(loops=0, sample_format=None, multi="sharded")
    for n in (1, 2, 4, 8):
        name = B.DEFAULT_BY_GPUS[n]
        W = B.prepare_workload(name, a, n)
        assert W["blocks_per_slot"] % n == 0 and W["blocks_per_slot"] >= B.WORKLOADS[name]["blocks_per_slot"]
        assert W["sfmt"] == "cs16" and W["multi"] == ("single" if n == 1 else "sharded")
        isz = O.geometry(W["sr"])[2].input_size
        P = dict(nblocks=W["blocks_per_slot"], nsamp=W["blocks_per_slot"] * isz)
        c1, c2 = B.config_of(name, W, P, n), B.config_of(name, B.prepare_workload(name, a, n), P, n)
        assert c1 == c2 and c1["sample_format"] == "CS16" and c1["channels"] == W["nch"]
    assert B.prepare_workload("cfg3", a, 3)["multi"] == "broadcast"          # 128 channels do not divide by 3
    assert B.prepare_workload("cfg2", a, 1)["sfmt"] == "cf32"                 # BASELINE names CF32 for config 2
    a.sample_format = "cf32"
    assert B.prepare_workload("cfg3", a, 1)["sfmt"] == "cf32"
