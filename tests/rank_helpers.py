"""Test helpers for the world_size-2 gloo tests (tests/test_multi_rank.py): the channel -> rank map the C ABI uses
(hfdl_b200_set_exchange: channel k of the job belongs to rank k % nranks), a capture broadcast and a PDU gather.
The multi-GPU data path itself is in the library (hfdl_b200_spectrum_slices / _slices_to / process_slices / push_peer,
block_shim.cu for ngpus > 1); bench.py drives it with one process per GPU."""


def shard_channels(freqs, rank, world):
    """Round-robin channel -> rank map (channel k on rank k % world); returns (indices, freqs) of this rank."""
    idx = list(range(rank, len(freqs), world))
    return idx, [freqs[i] for i in idx]


def broadcast_capture(tensor, src=0):
    """Make rank src's capture visible to every rank (no-op without an initialised process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src=src)
    return tensor


def gather_pdus(local):
    """PDUs go from every rank straight to the host side (no reduction): returns the canonical merged list
    [(sample_cnt_end, freq, octets, M1, crc_good), ...] on every rank."""
    import torch.distributed as dist
    items = [(int(q.sample_cnt_end), int(q.freq), q.data(), int(q.M1), int(q.crc_good)) for q in local]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, items)
        items = [x for part in out for x in part]
    return sorted(items)
