"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when built, the reference's own
sources compiled in place (oracle/_ref/libref.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

MAX_PDU = 945


class Ddc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("pre_decimation", "post_decimation", "taps_length", "taps_min_length",
                                         "overlap_length", "fft_size", "fft_inv_size", "input_size", "post_input_size")] + \
               [("pre_shift", C.c_float), ("startbin", C.c_int32), ("v", C.c_int32), ("offsetbin", C.c_int32),
                ("post_shift", C.c_float), ("scrap", C.c_int32),
                ("dsa_sindelta", C.c_float), ("dsa_cosdelta", C.c_float), ("dsa_rate", C.c_float)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Pdu(C.Structure):
    _fields_ = [("freq", C.c_int32), ("M1", C.c_int32), ("len", C.c_int32),
                ("freq_err_hz", C.c_float), ("signal_level", C.c_float), ("noise_floor", C.c_float),
                ("bit_rate", C.c_int32), ("slot", C.c_char),
                ("sample_cnt_end", C.c_uint64), ("sample_cnt_a2", C.c_uint64),
                ("train_bits_bad", C.c_int32), ("train_bits_total", C.c_int32), ("crc_good", C.c_int32),
                ("octets", C.c_uint8 * (MAX_PDU + 3))]

    def data(self):
        return bytes(self.octets[: self.len])


class TxFrame(C.Structure):
    _fields_ = [("freq_hz", C.c_int32), ("M1", C.c_int32), ("start_s", C.c_double), ("cfo_hz", C.c_double),
                ("phase0", C.c_double), ("amplitude", C.c_double), ("pdu_len", C.c_int32),
                ("pdu", C.c_uint8 * (MAX_PDU + 3))]


def build(fast=False):
    """(Re)build the oracle with make; cheap when up to date."""
    subprocess.run(["make", "-s", "-C", ODIR, "all"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_libs = {}


def lib(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(ODIR, name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    cfp = np.ctypeslib.ndpointer(np.complex64, flags="C")
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
    L.orc_fft.argtypes = [cfp, cfp, C.c_int, C.c_int]
    L.orc_fft_set_threads.argtypes = [C.c_int]
    L.orc_next_pow2.restype = C.c_int32
    L.orc_fft_decimation_rate.restype = C.c_int32
    L.orc_fft_decimation_rate.argtypes = [C.c_int32, C.c_int32]
    L.orc_relative_transition_bw.restype = C.c_float
    L.orc_relative_transition_bw.argtypes = [C.c_int32, C.c_int32]
    L.orc_ddc_init.argtypes = [C.POINTER(Ddc), C.c_float, C.c_int32, C.c_float]
    L.orc_channel_shift_rate.restype = C.c_float
    L.orc_channel_shift_rate.argtypes = [C.c_int32] * 3
    L.orc_bandpass_taps.argtypes = [cfp, C.c_int32, C.c_float, C.c_float]
    L.orc_channelizer_create.restype = C.c_void_p
    L.orc_channelizer_create.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_int]
    L.orc_channelizer_destroy.argtypes = [C.c_void_p]
    L.orc_channelizer_taps.restype = C.POINTER(C.c_float)
    L.orc_channelizer_taps.argtypes = [C.c_void_p]
    L.orc_channelizer_ddc.restype = C.POINTER(Ddc)
    L.orc_channelizer_ddc.argtypes = [C.c_void_p]
    L.orc_channelizer_execute.argtypes = [C.c_void_p, cfp, cfp]
    L.orc_swap_sides.argtypes = [cfp, C.c_int32]
    L.orc_firdes_kaiser.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, np.ctypeslib.ndpointer(np.float32)]
    L.orc_resamp_create.restype = C.c_void_p
    L.orc_resamp_create.argtypes = [C.c_float, C.c_float]
    L.orc_resamp_destroy.argtypes = [C.c_void_p]
    L.orc_resamp_execute.argtypes = [C.c_void_p, cfp, C.c_int, cfp, C.POINTER(C.c_uint32)]
    L.orc_resamp_design.argtypes = [C.c_float, C.c_float, np.ctypeslib.ndpointer(np.float32), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
    L.orc_crc16.restype = C.c_uint16
    L.orc_crc16.argtypes = [u8p, C.c_uint32, C.c_uint16]
    L.orc_pdu_crc_good.argtypes = [u8p, C.c_uint32]
    L.orc_viterbi27.argtypes = [u8p, C.c_int, u8p]
    L.orc_pdu_front_parse.argtypes = [u8p, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
    L.orc_pdu_front_parse.restype = None
    L.orc_conv_encode27.argtypes = [u8p, C.c_int, u8p]
    L.orc_scrambler_bits.argtypes = [u8p, C.c_int]
    L.orc_scrambler_bits.restype = C.c_uint32
    L.orc_decode_user_data.argtypes = [cfp, C.c_int, C.c_uint32, u8p, C.c_void_p]
    L.orc_encode_user_data.argtypes = [u8p, C.c_int, cfp]
    L.orc_pdu_len_octets.argtypes = [C.c_int]
    L.orc_channel_create.restype = C.c_void_p
    L.orc_channel_create.argtypes = [C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int]
    L.orc_channel_destroy.argtypes = [C.c_void_p]
    L.orc_channel_set_capture.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t]
    L.orc_channel_get_capture.restype = C.c_size_t
    L.orc_channel_get_capture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.orc_channel_process_block.argtypes = [C.c_void_p, cfp]
    L.orc_channel_process_baseband.argtypes = [C.c_void_p, cfp, C.c_int]
    L.orc_channel_pdu_count.argtypes = [C.c_void_p]
    L.orc_channel_get_pdu.argtypes = [C.c_void_p, C.c_int, C.POINTER(Pdu)]
    L.orc_channel_ddc.restype = C.POINTER(Ddc)
    L.orc_channel_ddc.argtypes = [C.c_void_p]
    L.orc_channel_resamp_rate.restype = C.c_float
    L.orc_channel_resamp_rate.argtypes = [C.c_void_p]
    L.orc_channel_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 4
    L.orc_channel_m1_not_found.argtypes = [C.c_void_p]
    L.orc_channel_noise_floor.argtypes = [C.c_void_p]
    L.orc_channel_noise_floor.restype = C.c_float
    L.orc_pipeline_create.restype = C.c_void_p
    L.orc_pipeline_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int, C.c_int]
    L.orc_pipeline_destroy.argtypes = [C.c_void_p]
    L.orc_pipeline_use_ring.argtypes = [C.c_void_p, C.c_int]
    L.orc_pipeline_sync.argtypes = [C.c_void_p]
    L.orc_pipeline_channel.restype = C.c_void_p
    L.orc_pipeline_channel.argtypes = [C.c_void_p, C.c_int]
    L.orc_pipeline_ddc.restype = C.POINTER(Ddc)
    L.orc_pipeline_ddc.argtypes = [C.c_void_p]
    L.orc_pipeline_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    L.orc_pipeline_pdu_count.argtypes = [C.c_void_p]
    L.orc_pipeline_get_pdu.argtypes = [C.c_void_p, C.c_int, C.POINTER(Pdu)]
    L.orc_pipeline_last_spectrum.argtypes = [C.c_void_p, cfp, C.c_int]
    L.orc_convert_samples.argtypes = [C.c_void_p, C.c_int64, C.c_int, cfp]
    L.orc_tx_make_pdu.argtypes = [C.c_int, C.c_int, C.c_uint64, u8p]
    L.orc_tx_frame_baseband.argtypes = [C.POINTER(TxFrame), cfp, C.c_int]
    L.orc_tx_frame_symbols.argtypes = [C.POINTER(TxFrame), cfp, C.c_int]
    L.orc_tx_render.argtypes = [cfp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(TxFrame), C.c_int, C.c_int, C.c_int]
    L.orc_tx_render_range.argtypes = [cfp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(TxFrame), C.c_int, C.c_int, C.c_int]
    L.orc_tx_add_noise.argtypes = [cfp, C.c_int64, C.c_double, C.c_uint64, C.c_int]
    L.orc_quantize_cs16.argtypes = [cfp, C.c_int64, np.ctypeslib.ndpointer(np.int16)]
    L.orc_quantize_cu8.argtypes = [cfp, C.c_int64, np.ctypeslib.ndpointer(np.uint8)]
    _libs[name] = L
    return L


def reflib():
    """The reference's own fastddc.c/libcsdr*.c/crc.c/viterbi27_port.c (None if not built)."""
    if "ref" in _libs:
        return _libs["ref"]
    path = os.path.join(ODIR, "_ref", "libref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src"):
            build()
        if not os.path.exists(path):
            _libs["ref"] = None
            return None
    R = C.CDLL(path)
    cfp = np.ctypeslib.ndpointer(np.complex64, flags="C")
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
    R.crc16_ccitt.restype = C.c_uint16
    R.crc16_ccitt.argtypes = [u8p, C.c_uint32, C.c_uint16]
    R.create_viterbi27.restype = C.c_void_p
    R.create_viterbi27.argtypes = [C.c_int]
    R.init_viterbi27.argtypes = [C.c_void_p, C.c_int]
    R.update_viterbi27_blk.argtypes = [C.c_void_p, u8p, C.c_int]
    R.chainback_viterbi27.argtypes = [C.c_void_p, u8p, C.c_uint, C.c_uint]
    R.delete_viterbi27.argtypes = [C.c_void_p]
    R.next_pow2.restype = C.c_int32
    R.next_pow2.argtypes = [C.c_int32]
    R.compute_fft_decimation_rate.restype = C.c_int32
    R.compute_fft_decimation_rate.argtypes = [C.c_int32, C.c_int32]
    R.compute_filter_relative_transition_bw.restype = C.c_float
    R.compute_filter_relative_transition_bw.argtypes = [C.c_int32, C.c_int32]
    R.fastddc_init.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.c_float]
    R.ref_sizeof_fastddc.restype = C.c_int
    R.ref_fastddc_fields.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    R.firdes_bandpass_c.argtypes = [cfp, C.c_int32, C.c_float, C.c_float, C.c_int]
    R.fft_channelizer_create.restype = C.c_void_p
    R.fft_channelizer_create.argtypes = [C.c_int32, C.c_float, C.c_float]
    R.fft_channelizer_destroy.argtypes = [C.c_void_p]
    R.ref_channelizer_taps_fft.restype = C.POINTER(C.c_float)
    R.ref_channelizer_taps_fft.argtypes = [C.c_void_p]
    R.ref_channelizer_ddc.restype = C.c_void_p
    R.ref_channelizer_ddc.argtypes = [C.c_void_p]
    R.ref_channelizer_execute.argtypes = [C.c_void_p, cfp, cfp]
    R.fft_swap_sides.argtypes = [cfp, C.c_int32]
    _bind_ref_host(R)
    _libs["ref"] = R
    return R


class RefPdu(C.Structure):
    """capture record of oracle/ref_shim/ref_host.c: struct hfdl_pdu_metadata (pdu.h:8-17) + the octet string"""
    _fields_ = [("version", C.c_int32), ("freq", C.c_int32), ("bit_rate", C.c_int32),
                ("freq_err_hz", C.c_float), ("rssi", C.c_float), ("noise_floor", C.c_float),
                ("slot", C.c_char), ("len", C.c_int32), ("flags", C.c_uint32), ("octets", C.c_uint8 * 948)]

    def data(self):
        return bytes(self.octets[: self.len])


def _bind_ref_host(R):
    if not hasattr(R, "ref_pipeline_create"):
        return
    R.ref_pipeline_create.restype = C.c_void_p
    R.ref_pipeline_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32]
    R.ref_pipeline_feed.restype = C.c_int64
    R.ref_pipeline_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    R.ref_pipeline_finish.argtypes = [C.c_void_p]
    R.ref_pipeline_drain.argtypes = [C.c_void_p]
    R.ref_pipeline_destroy.argtypes = [C.c_void_p]
    R.ref_pipeline_input_size.argtypes = [C.c_void_p]
    R.ref_pdu_get.argtypes = [C.c_int, C.POINTER(RefPdu)]
    R.ref_stat_count.restype = C.c_long
    R.ref_stat_count.argtypes = [C.c_int32, C.c_char_p]
    R.ref_dump_read.restype = C.c_long
    R.ref_dump_read.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long]
    R.ref_dumps_enable.argtypes = [C.c_int]
    R.ref_struct_sizes.argtypes = [C.c_int]


def reflib_fast():
    """Timing build of the same reference sources (-O3 -ffast-math, no DATADUMPS); None if not built."""
    if "ref_fast" in _libs:
        return _libs["ref_fast"]
    path = os.path.join(ODIR, "_ref", "libref_fast.so")
    if not os.path.exists(path) and os.path.isdir("/root/reference/src"):
        build()
    R = C.CDLL(path) if os.path.exists(path) else None
    if R is not None:
        _bind_ref_host(R)
        R.ref_fft_backend.restype = C.c_int
    _libs["ref_fast"] = R
    return R


def reflib_front():
    """The reference's own pdu.c / mpdu.c / spdu.c / lpdu.c / util.c / crc.c around its real pdu_decoder_thread
    (oracle/ref_shim/front/ref_front_host.c); None if not built."""
    if "ref_front" in _libs:
        return _libs["ref_front"]
    path = os.path.join(ODIR, "_ref", "libref_front.so")
    if not os.path.exists(path) and os.path.isdir("/root/reference/src"):
        build()
    R = C.CDLL(path) if os.path.exists(path) else None
    if R is not None:
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
        R.ref_front_counter_name.restype = C.c_char_p
        R.ref_front_run.argtypes = [u8p, np.ctypeslib.ndpointer(np.int32, flags="C"), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    np.ctypeslib.ndpointer(np.int64, flags="C")]
        R.ref_front_fcs_check.argtypes = [u8p, C.c_uint32]
        R.crc16_ccitt.restype = C.c_uint16
        R.crc16_ccitt.argtypes = [u8p, C.c_uint32, C.c_uint16]
    _libs["ref_front"] = R
    return R


def ref_front_run(pdus, freq=10063000, output_mpdus=False, output_corrupted=False):
    """Every PDU through the reference's pdu_decoder_thread, one at a time: list of {statsd counter / delivered-node name: count}."""
    R = reflib_front()
    K = R.ref_front_counter_count()
    names = [R.ref_front_counter_name(i).decode() for i in range(K)]
    buf = np.frombuffer(b"".join(bytes(p) for p in pdus) + b"\0", np.uint8).copy()
    lens = np.array([len(p) for p in pdus], np.int32)
    out = np.zeros((len(pdus), K), np.int64)
    r = R.ref_front_run(buf, lens, len(pdus), freq, int(output_mpdus), int(output_corrupted), out)
    assert r == 0, r
    return [dict(zip(names, (int(v) for v in row))) for row in out]


class RefPipeline:
    """The REFERENCE's own block.c + fft.c + fastddc.c + hfdl.c + viterbi27_port.c, wired as main.c does, fed as
    input-file.c does (oracle/ref_shim/ref_host.c).  One instance at a time (the PDU capture list is global)."""

    def __init__(self, sample_rate, centerfreq, freqs, sfmt=3, fft_threads=4, fast=False, dumps=False):
        self.R = reflib_fast() if fast else reflib()
        assert self.R is not None, "oracle/_ref not built"
        self.R.ref_pdu_clear()
        self.R.ref_stat_clear()
        self.R.ref_dumps_clear()
        self.R.ref_dumps_enable(1 if dumps else 0)
        fa = (C.c_int32 * len(freqs))(*freqs)
        self.p = self.R.ref_pipeline_create(sample_rate, centerfreq, fa, len(freqs), sfmt, fft_threads)
        assert self.p
        self.freqs = list(freqs)
        self.sfmt = sfmt
        self.finished = False

    def feed(self, raw):
        raw = np.ascontiguousarray(raw)
        n = raw.size if raw.dtype == np.complex64 else raw.size // 2
        return self.R.ref_pipeline_feed(self.p, raw.ctypes.data, n)

    def drain(self):
        self.R.ref_pipeline_drain(self.p)

    def finish(self):
        if not self.finished:
            self.R.ref_pipeline_finish(self.p)
            self.finished = True

    def pdus(self):
        out = []
        for i in range(self.R.ref_pdu_count()):
            q = RefPdu()
            self.R.ref_pdu_get(i, C.byref(q))
            out.append(q)
        return out

    def stat(self, freq, name):
        return int(self.R.ref_stat_count(freq, name.encode()))

    def dump(self, name, idx=0, complex_=True):
        n = self.R.ref_dump_read(name.encode(), idx, None, None, 0)
        if n < 0:
            return None, None
        t = np.zeros(n, np.uint64)
        v = np.zeros(n, np.complex64)
        self.R.ref_dump_read(name.encode(), idx, t.ctypes.data, v.ctypes.data, n)
        return t, (v if complex_ else v.real.copy())

    def close(self):
        if self.p:
            self.finish()
            self.R.ref_pipeline_destroy(self.p)
            self.p = None


# ---------------------------------------------------------------- helpers
FOLD_FULL, FOLD_SLICE = 0, 1
SFMT_CU8, SFMT_CS16, SFMT_CF32 = 1, 2, 3
CAP = dict(chan=0, agc=1, mf=2, symsync=3, costas=4, eq=5, datasym=6, ddc=7)


def ddc_init(tbw, decimation, shift_rate):
    d = Ddc()
    rc = lib().orc_ddc_init(C.byref(d), tbw, decimation, shift_rate)
    return d, rc


def geometry(sample_rate):
    L = lib()
    dec = L.orc_fft_decimation_rate(sample_rate, 5400)
    tbw = L.orc_relative_transition_bw(sample_rate, 250)
    d, _ = ddc_init(tbw, dec, 0.0)
    return dec, tbw, d


def pdu_front(buf):
    """orc_pdu_front_parse -> (status, direction, lpdus processed, good, bad_fcs, too_short, good mask)"""
    a = np.frombuffer(bytes(buf), np.uint8).copy()
    out = (C.c_int32 * 6)()
    m = C.c_uint64()
    lib().orc_pdu_front_parse(a, a.size, out, C.byref(m))
    return tuple(out) + (m.value,)


def slice_taps(sample_rate, centerfreq, freq):
    """tap-spectrum slice of one channel as the oracle's slice fold uses it: M bins, index i <-> swapped bin startbin-M/2+i"""
    L = lib()
    dec, tbw, _ = geometry(sample_rate)
    c = L.orc_channelizer_create(dec, tbw, L.orc_channel_shift_rate(sample_rate, centerfreq, freq), FOLD_SLICE)
    M = L.orc_channelizer_ddc(c).contents.fft_inv_size
    out = np.ctypeslib.as_array(L.orc_channelizer_taps(c), shape=(2 * M,)).copy().view(np.complex64)
    L.orc_channelizer_destroy(c)
    return out


def make_pdu(M1, kind=0, seed=1):
    L = lib()
    buf = np.zeros(MAX_PDU + 3, np.uint8)
    n = L.orc_tx_make_pdu(M1, kind, seed, buf)
    return bytes(buf[:n])


def noise_sigma(amplitude, sample_rate, esn0_db):
    """AWGN sigma per real component giving Es/N0 (dB) for a frame of linear 'amplitude':
    P = 3*sum(h^2)*amp^2 = 0.947 amp^2 (hfdl.c:148-154 taps at 3 sps), Es = P/1800, N0 = 2 sigma^2 / sample_rate."""
    P = 0.947 * amplitude * amplitude
    return float(np.sqrt(P * sample_rate / (1800 * 2 * 10 ** (esn0_db / 10))))


def tx_frame(freq_hz, M1, start_s, pdu, cfo_hz=0.0, phase0=0.0, amplitude=0.1):
    f = TxFrame()
    f.freq_hz, f.M1, f.start_s, f.cfo_hz, f.phase0, f.amplitude = freq_hz, M1, start_s, cfo_hz, phase0, amplitude
    f.pdu_len = len(pdu)
    for i, b in enumerate(pdu):
        f.pdu[i] = b
    return f


def render(nsamples, sample_rate, centerfreq, frames, noise_sigma=0.0, seed=1, cyclic=False, nthreads=8):
    L = lib()
    out = np.zeros(nsamples, np.complex64)
    arr = (TxFrame * len(frames))(*frames)
    L.orc_tx_render(out, nsamples, sample_rate, centerfreq, arr, len(frames), int(cyclic), nthreads)
    if noise_sigma > 0:
        L.orc_tx_add_noise(out, nsamples, noise_sigma, seed, nthreads)
    return out


def render_range(first, count, nsamples, sample_rate, centerfreq, frames, noise_sigma=0.0, seed=1, cyclic=False, nthreads=8):
    """samples [first, first + count) of the capture render() would produce (noise stream seeded per range)"""
    L = lib()
    out = np.zeros(count, np.complex64)
    arr = (TxFrame * len(frames))(*frames)
    L.orc_tx_render_range(out, first, count, nsamples, sample_rate, centerfreq, arr, len(frames), int(cyclic), nthreads)
    if noise_sigma > 0:
        L.orc_tx_add_noise(out, count, noise_sigma, seed, nthreads)
    return out


class Pipeline:
    def __init__(self, sample_rate, centerfreq, freqs, fold_mode=FOLD_FULL, nthreads=8, fast=False, ring_depth=0):
        self.L = lib(fast)
        fa = (C.c_int32 * len(freqs))(*freqs)
        self.p = self.L.orc_pipeline_create(sample_rate, centerfreq, fa, len(freqs), fold_mode, nthreads)
        assert self.p
        if ring_depth > 0:       # spectrum ring (include/hfdl_b200_ring.h) instead of the reference's barrier pair
            assert self.L.orc_pipeline_use_ring(self.p, ring_depth) == 0
        self.freqs = list(freqs)
        self.ddc = self.L.orc_pipeline_ddc(self.p).contents

    def feed(self, raw, sfmt=SFMT_CF32):
        raw = np.ascontiguousarray(raw)
        n = raw.size if sfmt == SFMT_CF32 and raw.dtype == np.complex64 else raw.size // 2
        return self.L.orc_pipeline_feed(self.p, raw.ctypes.data, n, sfmt)

    def sync(self):
        self.L.orc_pipeline_sync(self.p)

    def pdus(self):
        n = self.L.orc_pipeline_pdu_count(self.p)
        out = []
        for i in range(n):
            q = Pdu()
            self.L.orc_pipeline_get_pdu(self.p, i, C.byref(q))
            out.append(q)
        return out

    def channel(self, i):
        return self.L.orc_pipeline_channel(self.p, i)

    def set_capture(self, ch, taps, maxn):
        mask = 0
        for t in taps:
            mask |= 1 << CAP[t]
        self.L.orc_channel_set_capture(self.channel(ch), mask, maxn)

    def capture(self, ch, tap):
        c = self.channel(ch)
        n = self.L.orc_channel_get_capture(c, CAP[tap], None, 0)
        out = np.zeros(n, np.complex64)
        self.L.orc_channel_get_capture(c, CAP[tap], out.ctypes.data, n)
        return out

    def stats(self, ch):
        v = [C.c_int32() for _ in range(4)]
        self.L.orc_channel_stats(self.channel(ch), *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def m1_not_found(self, ch):
        return int(self.L.orc_channel_m1_not_found(self.channel(ch)))

    def noise_floor(self, ch):
        return float(self.L.orc_channel_noise_floor(self.channel(ch)))

    def last_spectrum(self):
        out = np.zeros(self.ddc.fft_size, np.complex64)
        self.L.orc_pipeline_last_spectrum(self.p, out, out.size)
        return out

    def close(self):
        if self.p:
            self.L.orc_pipeline_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
