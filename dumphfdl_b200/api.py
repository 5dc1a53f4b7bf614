"""ctypes bindings of include/hfdl_b200.h (no arithmetic here; every call lands in libhfdl_b200.so)."""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhfdl_b200.so")
SFMT_CU8, SFMT_CS16, SFMT_CF32 = 1, 2, 3
MAX_PDU = 945
CP = dict(spectrum=0, ddc=1, chan=2, agc=3, mf=4, eq=5, tapslice=6)


class Config(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("centerfreq_hz", C.c_int32), ("freqs_hz", C.POINTER(C.c_int32)),
                ("nfreq", C.c_int32), ("sample_format", C.c_int32), ("device", C.c_int32),
                ("max_blocks_per_batch", C.c_int32), ("capture_channel", C.c_int32), ("capture_max", C.c_int32)]


class Geometry(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("decimation", "pre_decimation", "post_decimation", "taps_length", "overlap_length",
                                         "fft_size", "fft_inv_size", "input_size", "post_input_size", "scrap", "out_per_block")] + \
               [("transition_bw", C.c_float), ("resamp_rate", C.c_float), ("fft_passes", C.c_int32), ("fft_len", C.c_int32 * 3)]


class Counters(C.Structure):
    """hfdl_b200_counters_t: the reference's per-channel statsd metrics (doc/STATSD_METRICS.md)"""
    _fields_ = [("freq", C.c_int32)] + [(n, C.c_int64) for n in (
        "A1_found", "A2_found", "M1_found", "M1_not_found", "frames_processed", "frames_good", "frames_bad_fcs", "frames_too_short",
        "frames_air2gnd", "frames_gnd2air", "lpdus_processed", "lpdus_good", "lpdus_bad_fcs", "lpdus_too_short")] + [("noise_floor", C.c_float)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Pdu(C.Structure):
    _fields_ = [("version", C.c_int32), ("freq", C.c_int32), ("bit_rate", C.c_int32), ("freq_err_hz", C.c_float),
                ("rssi", C.c_float), ("noise_floor", C.c_float), ("slot", C.c_char), ("M1", C.c_int32), ("crc_good", C.c_int32),
                ("train_bits_bad", C.c_int32), ("train_bits_total", C.c_int32), ("sample_cnt_a2", C.c_uint64),
                ("sample_cnt_end", C.c_uint64), ("rx_time_s", C.c_double), ("signal_level", C.c_float),
                ("noise_floor_lin", C.c_float), ("len", C.c_int32), ("octets", C.c_uint8 * (MAX_PDU + 3)),
                ("frame_status", C.c_int32), ("direction", C.c_int32), ("lpdus_processed", C.c_int32), ("lpdus_good", C.c_int32),
                ("lpdus_bad_fcs", C.c_int32), ("lpdus_too_short", C.c_int32), ("lpdu_good_mask", C.c_uint64)]

    def data(self):
        return bytes(self.octets[: self.len])


_lib = None


def bind(L):
    """Declare the prototypes of include/hfdl_b200.h on an already opened library handle."""
    vp = C.c_void_p
    L.hfdl_b200_device_count.restype = C.c_int32
    L.hfdl_b200_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.hfdl_b200_destroy.argtypes = [vp]
    L.hfdl_b200_destroy.restype = None
    L.hfdl_b200_get_geometry.argtypes = [vp, C.POINTER(Geometry)]
    L.hfdl_b200_push_samples.argtypes = [vp, vp, C.c_int64]
    L.hfdl_b200_push_samples_nowait.argtypes = [vp, vp, C.c_int64]
    L.hfdl_b200_wait_host_buffer.argtypes = [vp]
    L.hfdl_b200_pop_pdus.argtypes = [vp, vp, C.c_int32]
    L.hfdl_b200_flush.argtypes = [vp]
    L.hfdl_b200_process_device.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int32]
    L.hfdl_b200_sync.argtypes = [vp]
    L.hfdl_b200_submit.argtypes = [vp]
    L.hfdl_b200_wait_input.argtypes = [vp, C.c_int32]
    L.hfdl_b200_push_peer.argtypes = [vp, vp]
    L.hfdl_b200_poll.argtypes = [vp]
    L.hfdl_b200_set_exchange.argtypes = [vp, C.POINTER(C.c_int32), C.c_int32, C.c_int32]
    L.hfdl_b200_spectrum_slices.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp]
    L.hfdl_b200_spectrum_slices_to.argtypes = [vp, vp, C.c_int64, C.c_int32, C.POINTER(vp), C.c_int32, vp]
    L.hfdl_b200_process_slices.argtypes = [vp, vp, C.c_int32, vp]
    L.hfdl_b200_slice_elems.argtypes = [vp]
    L.hfdl_b200_slice_elems.restype = C.c_int64
    L.hfdl_b200_busy.argtypes = [vp]
    L.hfdl_b200_channel_counters.argtypes = [vp, C.c_int32, C.POINTER(Counters)]
    L.hfdl_b200_pdu_count.argtypes = [vp]
    L.hfdl_b200_pop_pdu.argtypes = [vp, C.POINTER(Pdu)]
    L.hfdl_b200_channel_noise_floor.argtypes = [vp, C.c_int32, C.POINTER(C.c_float)]
    L.hfdl_b200_channel_stats.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32)]
    L.hfdl_b200_print_summary.argtypes = [vp]
    L.hfdl_b200_print_summary.restype = None
    L.hfdl_b200_timer_start.argtypes = [vp]
    L.hfdl_b200_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.hfdl_b200_profile_enable.argtypes = [vp, C.c_int32]
    L.hfdl_b200_profile_read.argtypes = [vp, C.c_int32, vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    L.hfdl_b200_kernel_launches.argtypes = [vp]
    L.hfdl_b200_kernel_launches.restype = C.c_int64
    L.hfdl_b200_result_bytes_per_batch.argtypes = [vp]
    L.hfdl_b200_result_bytes_per_batch.restype = C.c_int64
    L.hfdl_b200_read_checkpoint.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int64]
    L.hfdl_b200_read_checkpoint.restype = C.c_int64
    L.hfdl_b200_fft_forward.argtypes = [C.c_int32, vp, vp, C.c_int32, C.c_int32]
    L.hfdl_b200_fec_decode.argtypes = [C.c_int32, vp, C.c_int32, C.c_int32, C.c_uint32, vp, C.c_int32, vp, vp]
    L.hfdl_b200_viterbi27.argtypes = [C.c_int32, vp, C.c_int32, C.c_int32, vp]
    L.hfdl_b200_pdu_len.argtypes = [C.c_int32]
    L.hfdl_b200_pdu_front_parse.argtypes = [C.c_int32, vp, C.c_int32, vp, C.c_int32, vp]
    return L


def load(path=None):
    """Open the CUDA library.  Raises (never falls back) when it is missing."""
    global _lib
    if path is None and _lib is not None:
        return _lib
    p = path or os.environ.get("HFDL_B200_LIB") or LIB_PATH      # HFDL_B200_LIB: a variant build of the same CUDA library (kernel experiments)
    if not os.path.exists(p):
        raise RuntimeError("%s not found: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback." % p)
    L = bind(C.CDLL(p))
    if path is None:
        _lib = L
    return L


def pdu_len(M1, lib=None):
    return (lib or load()).hfdl_b200_pdu_len(M1)


def fft_forward(x, device=0, lib=None):
    L = lib or load()
    x = np.ascontiguousarray(x, np.complex64)
    batch, n = (1, x.size) if x.ndim == 1 else x.shape
    out = np.empty_like(x)
    rc = L.hfdl_b200_fft_forward(device, x.ctypes.data, out.ctypes.data, n, batch)
    if rc != 0:
        raise RuntimeError("hfdl_b200_fft_forward failed")
    return out


def fec_decode(symbols, M1, bitmask=0, device=0, want_soft=False, lib=None):
    L = lib or load()
    symbols = np.ascontiguousarray(symbols, np.complex64)
    if symbols.ndim == 1:
        symbols = symbols[None, :]
    nf = symbols.shape[0]
    plen = L.hfdl_b200_pdu_len(M1)
    out = np.zeros((nf, plen), np.uint8)
    crc = np.zeros(nf, np.int32)
    soft = np.zeros((nf, 15120), np.uint8) if want_soft else None
    rc = L.hfdl_b200_fec_decode(device, symbols.ctypes.data, nf, M1, bitmask, out.ctypes.data, plen,
                                soft.ctypes.data if want_soft else None, crc.ctypes.data)
    if rc != 0:
        raise RuntimeError("hfdl_b200_fec_decode failed")
    return out, crc, soft


def pdu_front_parse(pdus, device=0, lib=None):
    """list of bytes -> list of (frame_status, direction, lpdus processed, good, bad_fcs, too_short, good mask, crc_good)"""
    L = lib or load()
    n = len(pdus)
    stride = max(len(p) for p in pdus)
    buf = np.zeros((n, stride), np.uint8)
    lens = np.zeros(n, np.int32)
    for i, p in enumerate(pdus):
        buf[i, :len(p)] = np.frombuffer(bytes(p), np.uint8)
        lens[i] = len(p)
    out = (Pdu * n)()
    if L.hfdl_b200_pdu_front_parse(device, buf.ctypes.data, stride, lens.ctypes.data, n, out) != 0:
        raise RuntimeError("hfdl_b200_pdu_front_parse failed")
    return [(q.frame_status, q.direction, q.lpdus_processed, q.lpdus_good, q.lpdus_bad_fcs, q.lpdus_too_short, q.lpdu_good_mask, q.crc_good) for q in out]


def viterbi27(syms, nbits, device=0, lib=None):
    L = lib or load()
    syms = np.ascontiguousarray(syms, np.uint8)
    if syms.ndim == 1:
        syms = syms[None, :]
    nf = syms.shape[0]
    out = np.zeros((nf, (nbits + 7) // 8), np.uint8)
    rc = L.hfdl_b200_viterbi27(device, syms.ctypes.data, nf, nbits, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("hfdl_b200_viterbi27 failed")
    return out


class Frontend:
    """hfdl_b200_create .. hfdl_b200_destroy as an object."""

    def __init__(self, sample_rate, centerfreq_hz, freqs_hz, sample_format=SFMT_CF32, device=0, max_blocks_per_batch=0,
                 capture_channel=-1, capture_max=0, lib=None):
        self.L = lib or load()
        self._freqs = (C.c_int32 * len(freqs_hz))(*freqs_hz)
        cfg = Config(sample_rate, centerfreq_hz, self._freqs, len(freqs_hz), sample_format, device,
                     max_blocks_per_batch, capture_channel, capture_max)
        self.h = C.c_void_p()
        if self.L.hfdl_b200_create(C.byref(self.h), C.byref(cfg)) != 0:
            raise RuntimeError("hfdl_b200_create failed")
        self.freqs = list(freqs_hz)
        self.sample_format = sample_format
        self.geom = Geometry()
        self.L.hfdl_b200_get_geometry(self.h, C.byref(self.geom))

    def push(self, samples):
        a = np.ascontiguousarray(samples)
        bps = {SFMT_CU8: 2, SFMT_CS16: 4, SFMT_CF32: 8}[self.sample_format]
        n = a.nbytes // bps
        r = self.L.hfdl_b200_push_samples(self.h, a.ctypes.data, n)
        if r < 0:
            raise RuntimeError("hfdl_b200_push_samples failed")
        return r

    def push_ptr(self, ptr, nsamples, wait=True):
        r = (self.L.hfdl_b200_push_samples if wait else self.L.hfdl_b200_push_samples_nowait)(self.h, ptr, nsamples)
        if r < 0:
            raise RuntimeError("hfdl_b200_push_samples failed")
        return r

    def flush(self):
        r = self.L.hfdl_b200_flush(self.h)
        if r < 0:
            raise RuntimeError("hfdl_b200_flush failed")
        return r

    def process_device(self, dptr, ring_samples, start_sample, nblocks):
        r = self.L.hfdl_b200_process_device(self.h, dptr, ring_samples, start_sample, nblocks)
        if r < 0:
            raise RuntimeError("hfdl_b200_process_device failed")
        return r

    def set_exchange(self, all_freqs_hz, nranks):
        arr = (C.c_int32 * len(all_freqs_hz))(*all_freqs_hz)
        if self.L.hfdl_b200_set_exchange(self.h, arr, len(all_freqs_hz), nranks) != 0:
            raise RuntimeError("hfdl_b200_set_exchange failed")

    def spectrum_slices(self, d_samples, first_block, nblocks, d_send, stream=None):
        r = self.L.hfdl_b200_spectrum_slices(self.h, d_samples, first_block, nblocks, d_send, stream)
        if r < 0:
            raise RuntimeError("hfdl_b200_spectrum_slices failed")
        return r

    def spectrum_slices_to(self, d_samples, first_block, nblocks, recv_ptrs, batch_block0, stream=None):
        arr = (C.c_void_p * len(recv_ptrs))(*recv_ptrs)
        r = self.L.hfdl_b200_spectrum_slices_to(self.h, d_samples, first_block, nblocks, arr, batch_block0, stream)
        if r < 0:
            raise RuntimeError("hfdl_b200_spectrum_slices_to failed")
        return r

    def process_slices(self, d_slices, nblocks, stream=None):
        r = self.L.hfdl_b200_process_slices(self.h, d_slices, nblocks, stream)
        if r < 0:
            raise RuntimeError("hfdl_b200_process_slices failed")
        return r

    def sync(self):
        self.L.hfdl_b200_sync(self.h)

    def push_peer(self, src):
        r = self.L.hfdl_b200_push_peer(self.h, src.h)
        if r < 0:
            raise RuntimeError("hfdl_b200_push_peer failed")
        return r

    def wait_input(self, keep=0):
        if self.L.hfdl_b200_wait_input(self.h, keep) != 0:
            raise RuntimeError("hfdl_b200_wait_input failed")

    def submit(self):
        r = self.L.hfdl_b200_submit(self.h)
        if r < 0:
            raise RuntimeError("hfdl_b200_submit failed")
        return r

    def poll(self):
        return self.L.hfdl_b200_poll(self.h)

    def busy(self):
        return self.L.hfdl_b200_busy(self.h)

    def counters(self, ch):
        c = Counters()
        if self.L.hfdl_b200_channel_counters(self.h, ch, C.byref(c)) != 0:
            raise RuntimeError("hfdl_b200_channel_counters failed")
        return c

    def wait_host_buffer(self):
        if self.L.hfdl_b200_wait_host_buffer(self.h) != 0:
            raise RuntimeError("hfdl_b200_wait_host_buffer failed")

    def pdus(self):
        out = []
        while True:
            n = self.L.hfdl_b200_pdu_count(self.h)
            if n <= 0:
                break
            arr = (Pdu * n)()
            k = self.L.hfdl_b200_pop_pdus(self.h, arr, n)
            out.extend(arr[i] for i in range(k))
            if k < n:
                break
        return out

    def stats(self, ch):
        v = (C.c_int32 * 4)()
        self.L.hfdl_b200_channel_stats(self.h, ch, v)
        return tuple(v)

    def noise_floor(self, ch):
        f = C.c_float()
        self.L.hfdl_b200_channel_noise_floor(self.h, ch, C.byref(f))
        return f.value

    def checkpoint(self, what, index=0):
        n = self.L.hfdl_b200_read_checkpoint(self.h, CP[what], index, None, 0)
        if n < 0:
            raise RuntimeError("checkpoint %s unavailable" % what)
        out = np.zeros(n, np.complex64)
        if n:
            self.L.hfdl_b200_read_checkpoint(self.h, CP[what], index, out.ctypes.data, n)
        return out

    def timer_start(self):
        self.L.hfdl_b200_timer_start(self.h)

    def timer_stop(self):
        ms = C.c_float()
        self.L.hfdl_b200_timer_stop(self.h, C.byref(ms))
        return ms.value

    def profile(self, on):
        self.L.hfdl_b200_profile_enable(self.h, int(on))

    def profile_read(self):
        names = ((C.c_char * 32) * 16)()
        ms = (C.c_float * 16)()
        n = (C.c_int32 * 16)()
        k = self.L.hfdl_b200_profile_read(self.h, 16, names, ms, n)
        return {names[i].value.decode(): (ms[i], n[i]) for i in range(k)}

    def launches(self):
        return self.L.hfdl_b200_kernel_launches(self.h)

    def result_bytes_per_batch(self):
        return self.L.hfdl_b200_result_bytes_per_batch(self.h)

    def close(self):
        if self.h:
            self.L.hfdl_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
