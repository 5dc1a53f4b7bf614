"""world_size-2 test of the N>1 path on CPU (gloo): rank 0's capture is broadcast, each rank runs its channel
shard through the front-end (host-emulation build of the kernels), PDUs are gathered and must equal the
oracle's for the full channel set."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

import b200_cases as K
import orclib as O

HERE = os.path.dirname(os.path.abspath(__file__))
FREQS = [10063000, 9952000, 10101000]
MODES = [1, 2, 0]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import dumphfdl_b200.api as A
    import rank_helpers as sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = A.bind(C.CDLL(os.path.join(HERE, "cusim", "libhfdl_cusim.so")))
    n = int(250000 * 3.3)
    if rank == 0:
        x, _ = K.make_capture(250000, FREQS, MODES, 3.3, seed=41)
        t = torch.from_numpy(x.view(np.float32).copy())
    else:
        t = torch.zeros(2 * n, dtype=torch.float32)
    sharding.broadcast_capture(t, 0)
    idx, mine = sharding.shard_channels(FREQS, rank, world)
    fe = A.Frontend(250000, K.CF, mine, max_blocks_per_batch=8, lib=sim)
    fe.push(t.numpy())
    fe.flush()
    merged = sharding.gather_pdus(fe.pdus())
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_channel_sharding_gloo():
    import torch.multiprocessing as mp
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim"), "all"], check=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x, truth = K.make_capture(250000, FREQS, MODES, 3.3, seed=41)
    ref = K.run_oracle(250000, FREQS, x, O.SFMT_CF32).pdus()
    want = sorted((int(r.sample_cnt_end), int(r.freq), r.data(), int(r.M1), int(r.crc_good)) for r in ref)
    assert merged == want and len(merged) == 3


def test_shard_map_is_a_partition():
    import rank_helpers as sharding
    f = list(range(1000, 1013))
    for world in (1, 2, 4, 8):
        parts = [sharding.shard_channels(f, r, world) for r in range(world)]
        assert sorted(i for idx, _ in parts for i in idx) == list(range(len(f)))
        assert all(fr == [f[i] for i in idx] for idx, fr in parts)


def test_render_range_slices_equal_the_whole_capture():
    """bench.py lets every rank render its own 1/N time slice of the looped slab (orc_tx_render_range): the slices put
    together must be the capture a single render produces, cyclic wrap included."""
    sr, n = 250000, 250000 * 3
    pd = O.make_pdu(1, 0, 5)
    frames = [O.tx_frame(FREQS[0], 1, 2.2, pd, cfo_hz=3.0, phase0=0.4, amplitude=0.1),       # wraps around the end of the slab
              O.tx_frame(FREQS[1], 2, 0.1, O.make_pdu(2, 1, 6), cfo_hz=-7.0, amplitude=0.07)]
    whole = O.render(n, sr, K.CF, frames, cyclic=True, nthreads=3)
    for world in (2, 3, 8):
        part = n // world
        pieces = [O.render_range(r * part, part if r < world - 1 else n - r * part, n, sr, K.CF, frames, cyclic=True, nthreads=2) for r in range(world)]
        assert np.array_equal(np.concatenate(pieces), whole)


def _worker_gather(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import dumphfdl_b200.api as A
    import rank_helpers as sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = A.bind(C.CDLL(os.path.join(HERE, "cusim", "libhfdl_cusim.so")))
    sr = 250000
    n = int(sr * 3.3) // world * world
    pdus = [O.make_pdu(m, i % 2, seed=41000 + i) for i, m in enumerate(MODES)]
    frames = [O.tx_frame(f, m, 0.15 + 0.03 * i, pdus[i], cfo_hz=4.0 * i - 3, phase0=0.5 * i, amplitude=0.1) for i, (f, m) in enumerate(zip(FREQS, MODES))]
    part = n // world
    # the host scatters the capture: this rank holds (and would upload) only its own time slice ...
    mine = O.render_range(rank * part, part, n, sr, K.CF, frames, noise_sigma=O.noise_sigma(0.1, sr, 20.0), seed=77 + rank, nthreads=2)
    t = torch.from_numpy(mine.view(np.float32).copy())
    # ... and an all-gather completes it on every rank (NCCL over NVLink on GPUs, gloo here)
    full = torch.zeros(2 * n, dtype=torch.float32)
    dist.all_gather_into_tensor(full, t)
    idx, myfreqs = sharding.shard_channels(FREQS, rank, world)
    fe = A.Frontend(sr, K.CF, myfreqs, max_blocks_per_batch=8, lib=sim)
    fe.push(full.numpy())
    fe.flush()
    merged = sharding.gather_pdus(fe.pdus())
    if rank == 0:
        q.put((merged, full.numpy().view(np.complex64).copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scatter_allgather_gloo():
    """The e2e path of bench.py at N > 1: every rank owns a time slice of the capture, an all-gather assembles it, the
    channels are sharded; the merged PDUs equal the oracle's on the assembled capture."""
    import torch.multiprocessing as mp
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim"), "all"], check=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_gather, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, x = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = K.run_oracle(250000, FREQS, x, O.SFMT_CF32).pdus()
    want = sorted((int(r.sample_cnt_end), int(r.freq), r.data(), int(r.M1), int(r.crc_good)) for r in ref)
    assert merged == want and len(merged) == 3


def _worker_sharded(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import dumphfdl_b200.api as A
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = A.bind(C.CDLL(os.path.join(HERE, "cusim", "libhfdl_cusim.so")))
    sr, batch = 250000, 4
    freqs = FREQS + [9931000]
    x, _ = K.make_capture(sr, freqs, MODES + [3], 3.3, seed=43)          # every rank renders the same capture, uses its share only
    fe = A.Frontend(sr, K.CF, freqs[rank::world], max_blocks_per_batch=batch, lib=sim)
    fe.set_exchange(freqs, world)
    g = fe.geom
    isz, ovl, M = g.input_size, g.overlap_length, g.fft_inv_size
    cper = len(freqs) // world
    nb = x.size // isz
    nb -= nb % world
    padded = np.concatenate([np.zeros(ovl, np.complex64), x[: nb * isz]])
    done, keep = 0, []
    while done < nb:
        B = min(batch, nb - done)
        bl = B // world
        first = done + rank * bl
        mine = np.ascontiguousarray(padded[first * isz: first * isz + ovl + bl * isz])      # what this rank would upload over its own PCIe link
        send = torch.zeros(world * bl * cper * M * 2, dtype=torch.float32)
        recv = torch.zeros_like(send)
        fe.spectrum_slices(mine.ctypes.data, first, bl, send.data_ptr())
        dist.all_to_all_single(recv, send)            # NCCL over NVLink on GPUs (bench.py), gloo here
        fe.process_slices(recv.data_ptr(), B)
        keep += [mine, send, recv]
        done += B
    fe.sync()
    import rank_helpers as sharding
    merged = sharding.gather_pdus(fe.pdus())
    if rank == 0:
        q.put((merged, nb * isz))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_spectrum_gloo():
    """bench.py's N > 1 data path: every rank transforms half of each batch's blocks for all channels, the pass-band
    slices change hands in an all-to-all, every rank demodulates its own channels; merged PDUs == the oracle's."""
    import torch.multiprocessing as mp
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "cusim"), "all"], check=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, used = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    freqs = FREQS + [9931000]
    x, truth = K.make_capture(250000, freqs, MODES + [3], 3.3, seed=43)
    ref = K.run_oracle(250000, freqs, x[:used], O.SFMT_CF32).pdus()
    want = sorted((int(r.sample_cnt_end), int(r.freq), r.data(), int(r.M1), int(r.crc_good)) for r in ref)
    assert merged == want and len(merged) == 4
