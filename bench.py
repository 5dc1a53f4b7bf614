#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (I/Q Msamples/s ingested, CRC-good PDUs/s) on BASELINE's own configurations.

  python bench.py --gpus N --steps K --warmup W            our arm  (libhfdl_b200.so, sm_100a kernels)
  python bench.py --impl reference --gpus N ...            the reference's CPU path on the host cores: the reference's
                                                           own block.c / fft.c / fastddc.c / hfdl.c / viterbi27_port.c
                                                           compiled where they lie (oracle/_ref/libref_fast.so, -O3
                                                           -ffast-math like src/CMakeLists.txt:39-42), one thread per
                                                           channel + FFT threads exactly as dumphfdl runs them; liquid-dsp
                                                           and fftw3f (not installed) are served by the oracle's objects /
                                                           FFT (fftw3f is used when the box has it)

Workload by GPU count (BASELINE.json configs): N = 1, 2 -> cfg3 (20 Msps, 128 channels: the largest single-GPU
configuration), N = 4 -> cfg4 (30 Msps, 256 channels), N = 8 -> cfg5 (60 Msps, 512 channels); --workload overrides
(cfg2 = 2 Msps, 8 channels is still available).  The capture is ONE looped multichannel slab (frames placed cyclically, so
the stream is seamless across passes); a "step" is LOOPS passes of the hot path over it.
  value : Msamples/s with the capture resident in HBM when the timed region starts (device timed, CUDA events).
          N > 1: resident on rank 0; every pass broadcasts it over NCCL/NVLink inside the timed region, every rank runs the
          forward FFT on the whole capture and demodulates its own channels (channel k on rank k mod N).
  e2e   : the same metric from pinned HOST memory, copies inside the timed region.  N = 1: hfdl_b200_push_samples (H2D
          inside the C-ABI call) + PDU records D2H.  N > 1: the host scatters the slab -- rank r uploads the r-th 1/N of
          every pass over its own PCIe link, an NCCL all-gather over NVLink completes the capture on every GPU -- then
          hfdl_b200_process_device; the capture crosses PCIe exactly once.
The synthetic capture and the list of transmitted PDUs come from the HFDL transmitter of the test infrastructure
(oracle/orc_tx.c via tests/orclib.py): input generation and the exactness check of the decoded PDUs, outside every timed
region and never on the product path.  The only oracle/ code that is TIMED is the cpu_baseline / --impl reference leg."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CF = 100000000            # centre 100 000 kHz (only differences matter)
ESN0_DB = 20.0
METRIC = "I/Q Msamples/s & CRC-good PDUs/s at 1/2/4/8 B200 vs fftw CPU ref"
# blocks_per_slot * input_size / sample_rate >= one 2.344 s single-slot frame; seed = 635000 + cfg index;
# loops = passes over the slab per step (so that K = 20 steps time more than a second of device work)
# sfmt: BASELINE names CF32 for cfg2 only.  The wideband configurations take CS16, the native format of the SDRs that deliver
# 20-60 Msps and the one dumphfdl's own SoapySDR input asks for first ("native sample format ... to avoid extra
# conversion", input-soapysdr.c:52-61); --sample-format overrides.
WORKLOADS = {
    "cfg2": dict(sr=2000000, nch=8, blocks_per_slot=22, slots=4, seed=635002, loops=16, group=0, sfmt="cf32"),
    "cfg3": dict(sr=20000000, nch=128, blocks_per_slot=14, slots=1, seed=635003, loops=32, group=0, sfmt="cs16"),
    "cfg4": dict(sr=30000000, nch=256, blocks_per_slot=21, slots=1, seed=635004, loops=16, group=64, sfmt="cs16"),
    "cfg5": dict(sr=60000000, nch=512, blocks_per_slot=21, slots=1, seed=635005, loops=16, group=64, sfmt="cs16"),
}
SFMT = {"cf32": dict(code=3, bps=8, np=np.float32, name="CF32"), "cs16": dict(code=2, bps=4, np=np.int16, name="CS16")}
DEFAULT_BY_GPUS = {1: "cfg3", 2: "cfg3", 4: "cfg4", 8: "cfg5"}


def channel_freqs(W):
    nch, sr = W["nch"], W["sr"]
    delta = int(0.85 * sr / nch / 1000) * 1000
    return [int(np.floor((CF + (k - (nch - 1) / 2) * delta) / 1000.0 + 0.5)) * 1000 for k in range(nch)], delta


def plan_frames(O, W, isz):
    """Cyclic slab: per channel, back-to-back frames of random modes filling the slots.  cfg4 / cfg5 (hundreds of
    channels at 30 / 60 Msps): frames are planned for the lowest `group` channels only; the other channels carry
    frequency-shifted copies of that sub-band (see render_range) -- same modulation, same PDUs, different frequency."""
    freqs, delta = channel_freqs(W)
    nbase = W["group"] or W["nch"]
    nblocks = W["blocks_per_slot"] * W["slots"]
    nsamp = nblocks * isz
    slot_s = W["blocks_per_slot"] * isz / W["sr"]
    rng = np.random.default_rng(W["seed"])
    amp = 0.25 / np.sqrt(W["nch"]) / np.sqrt(0.947)
    frames, base_truth = [], []
    for k in range(nbase):
        f = freqs[k]
        slot = 0
        phase = float(rng.uniform(0, slot_s))            # random start offset of this channel's slot grid
        while slot < W["slots"]:
            m = int(rng.integers(0, 8 if W["slots"] >= 2 else 4))
            need = 2 if m >= 4 else 1
            if slot + need > W["slots"]:
                m -= 4
                need = 1
            flen = (448 + 531 + (168 if m >= 4 else 72) * 45) / 1800.0 + 0.02
            start = (phase + slot * slot_s + float(rng.uniform(0, need * slot_s - flen))) % (nsamp / W["sr"])
            pdu = O.make_pdu(m, int(rng.integers(0, 2)), seed=int(rng.integers(1, 1 << 30)))
            frames.append(O.tx_frame(f, m, start, pdu, cfo_hz=float(rng.uniform(-20, 20)), phase0=float(rng.uniform(0, 2 * np.pi)), amplitude=amp))
            base_truth.append((k, pdu))
            slot += need
    ngroups = W["nch"] // nbase
    truth = [(freqs[k + g * nbase], pdu) for g in range(ngroups) for k, pdu in base_truth]
    return dict(freqs=freqs, delta=delta, frames=frames, truth=truth, nblocks=nblocks, nsamp=nsamp, amp=amp, nbase=nbase, ngroups=ngroups)


def render_range(O, W, P, first, count, nthreads, noise_seed):
    """Samples [first, first + count) of the looped slab."""
    sr = W["sr"]
    x = O.render_range(first, count, P["nsamp"], sr, CF, P["frames"], cyclic=True, nthreads=nthreads)
    if P["ngroups"] > 1:
        # sub-band copies: x(t) = sum_g base(t) * exp(j 2 pi F_g t), F_g = g * group * delta moved to the nearest multiple
        # of 1 / T_slab (< 0.2 Hz away) so that every copy is continuous across the wrap of the cyclic slab
        T = P["nsamp"] / sr
        n = np.arange(first, first + count, dtype=np.float64)
        out = x.copy()
        for g in range(1, P["ngroups"]):
            Fg = round(g * P["nbase"] * P["delta"] * T) / T
            ph = (Fg / sr * n) % 1.0
            out += x * np.exp(2j * np.pi * ph).astype(np.complex64)
        x = out
    sig = O.noise_sigma(P["amp"], sr, ESN0_DB)
    O.lib().orc_tx_add_noise(x, x.size, sig, noise_seed, nthreads)
    return x


def to_raw(O, W, x):
    """complex64 samples -> the workload's sample format (flat array of I/Q values)"""
    if W["sfmt"] == "cf32":
        return np.ascontiguousarray(x).view(np.float32)
    raw = np.zeros(2 * x.size, np.int16)
    O.lib().orc_quantize_cs16(np.ascontiguousarray(x), x.size, raw)
    return raw


def bind_near_gpu(gpu):
    """Pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host buffers it allocates (first
    touch) sit on the GPU's NUMA node and N ranks uploading at once do not all pull from one socket's memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < ncpu:
            os.sched_setaffinity(0, cpus)
            return "%d CPUs local to GPU %d" % (len(cpus), idx)
        return "no distinct CPU set for GPU %d" % idx
    except Exception as e:
        return "not bound (%s)" % type(e).__name__


def prepare_workload(name, a, world):
    """The workload both arms run: BASELINE's configuration for the GPU count, the sample format, and -- when the
    spectrum is sharded over `world` GPUs -- a slab whose block count divides by `world`."""
    W = dict(WORKLOADS[name])
    if a.loops > 0:
        W["loops"] = a.loops
    if a.sample_format:
        W["sfmt"] = a.sample_format
    W["multi"] = a.multi if world > 1 else "single"
    if world > 1 and W["nch"] % world:
        W["multi"] = "broadcast"                       # the sharded spectrum wants the same number of channels on every GPU
    if world > 1:
        W["blocks_per_slot"] = -(-W["blocks_per_slot"] // world) * world      # every rank renders / uploads / transforms whole blocks
    return W


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every 10 ms
    (nvidia_ml_py), nvidia-smi -lms as the fallback when NVML cannot be imported."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu):
        self.gpu = gpu
        self.sm, self.mx, self.mask = [], [], 0
        self.stop_flag = False
        self.t = None
        self.p = None
        self.rows = []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mx.append(mx)
                        try:
                            self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.t = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t2 = threading.Thread(target=self._read, daemon=True)
            self.t2.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def stop(self):
        if self.t is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            reasons = [n for n, bit in self.REASONS.items() if self.mask & bit]
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


def config_of(name, W, P, world):
    """Identical for both arms: names the workload only (arm-specific step sizes are reported under "step")."""
    if world > 1 and W["multi"] == "sharded":
        shard = ("%d GPUs, one capture: every GPU transforms 1/%d of each batch's overlap-save blocks for all channels (its share of the capture "
                 "comes over its own PCIe link, the overlap in front of it from the previous GPU over NVLink), the channels' pass-band spectrum slices change hands in an NCCL all-to-all over NVLink, channel k is "
                 "demodulated on GPU k mod %d (%d channels per GPU)" % (world, world, world, W["nch"] // world))
    elif world > 1:
        shard = "%d GPUs: one capture, channel k on GPU k mod %d (%d channels per GPU); value: NCCL broadcast from rank 0 every pass; e2e: host scatter (1/%d of every pass per PCIe link) + NCCL all-gather" % (world, world, W["nch"] // world, world)
    else:
        shard = "1 GPU: all %d channels" % W["nch"]
    F = SFMT[W["sfmt"]]
    return {"workload": "%s: %.0f Msps %s, %d HFDL channels, synthetic looped slab of %d overlap-save blocks (%.1f Msamples, %.0f MB > 126 MB L2, no explicit flush), "
                        "Es/N0 %.0f dB, seed %d" % (name, W["sr"] / 1e6, F["name"], W["nch"], P["nblocks"], P["nsamp"] / 1e6, P["nsamp"] * F["bps"] / 1e6, ESN0_DB, W["seed"]),
            "sample_rate": W["sr"], "sample_format": F["name"], "channels": W["nch"], "blocks_per_slab": P["nblocks"], "esn0_db": ESN0_DB, "sharding": shard}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_setup(O, W, P):
    cores = os.cpu_count() or 1
    fast = O.reflib_fast()
    if fast is not None:
        kind = "reference"
        fft_threads = max(1, min(cores, 8))          # --fft-threads (main.c:438, fft.h:15 default 4); the channel threads are one per channel
        p = O.RefPipeline(W["sr"], CF, P["freqs"], sfmt=SFMT[W["sfmt"]]["code"], fft_threads=fft_threads, fast=True)
        backend = {0: "oracle FFT (persistent pool) as the fftw3f stand-in", 1: "fftw3f", 3: "fftw3f + fftw3f_threads",
                   4: "cache-blocked four-step FFT on a persistent worker pool (oracle/ref_shim/fft4step.c) as the fftw3f + fftw3f_threads stand-in"}[fast.ref_fft_backend()]
        desc = ("the reference's own block.c + fft.c + fastddc.c (all-bin fold) + hfdl.c + libfec/viterbi27_port.c compiled where they lie "
                "(oracle/_ref/libref_fast.so, -O3 -ffast-math), wired as main.c does: 1 fft thread with %d FFT workers + %d channel threads on a %d-core host; "
                "FFT backend: %s; liquid-dsp objects served by the oracle's restatement (liquid-dsp not installed)" % (fft_threads, W["nch"], cores, backend))
        return dict(kind=kind, p=p, cores=min(cores, W["nch"] + fft_threads), desc=desc, feed=lambda seg: (p.feed(seg), p.drain()), close=p.close,
                    good=lambda: sum(1 for q in p.pdus() if O.pdu_front(q.data())[0] == 0))
    nt = min(cores, 32)
    p = O.Pipeline(W["sr"], CF, P["freqs"], fold_mode=O.FOLD_FULL, nthreads=nt, fast=True)
    desc = "restated CPU reference (oracle/liboracle_fast.so, -O3 -ffast-math, all-bin fold, %d threads on a %d-core host): oracle/_ref not built" % (nt, cores)
    return dict(kind="port", p=p, cores=nt, desc=desc, feed=lambda seg: p.feed(seg, SFMT[W["sfmt"]]["code"]), close=p.close, good=lambda: sum(1 for q in p.pdus() if q.crc_good))


def fft_share(O, W, isz, per_block_s):
    """Share of the forward FFT in one block of the CPU arm (timed alone on the same box)."""
    try:
        g = O.geometry(W["sr"])[2]
        n = g.fft_size
        x = (np.random.default_rng(1).standard_normal(n) + 0j).astype(np.complex64)
        o = np.zeros(n, np.complex64)
        R = O.reflib_fast()
        if R is not None and hasattr(R, "ref_fft_run"):          # the backend the reference arm really used (fftw3f or its stand-in)
            cfp = np.ctypeslib.ndpointer(np.complex64, flags="C")
            R.ref_fft_run.argtypes = [cfp, cfp, ctypes.c_int32, ctypes.c_int32]
            R.csdr_fft_init(max(1, min(os.cpu_count() or 1, 8)))          # as cpu_reference_setup asked for (no effect once the pool runs)
            run = lambda: R.ref_fft_run(x, o, n, 1)
        else:
            L = O.lib(True)
            L.orc_fft_set_threads(max(1, min(os.cpu_count() or 1, 8)))
            run = lambda: L.orc_fft(x, o, n, 1)
        run()
        t0 = time.perf_counter()
        for _ in range(3):
            run()
        return (time.perf_counter() - t0) / 3 / per_block_s
    except Exception:
        return None


def cpu_reference(O, W, P, isz, x0, target_s=12.0):
    """cpu_baseline of the N = 1 line: a bounded sample (~target_s of CPU work) of the same workload: the looped slab
    (x0 = the whole slab) fed block by block, cyclically, the way the GPU arm loops over it."""
    R = cpu_reference_setup(O, W, P)
    v = 2 * isz                              # x0: the slab in the workload's sample format, two values per sample
    navail = x0.size // v
    R["feed"](x0[: v])                       # warm up: page in, first block
    t0 = time.perf_counter()
    R["feed"](x0[v: 2 * v])
    per_block = time.perf_counter() - t0
    nb = int(max(2, min(20 * navail, target_s / max(per_block, 1e-6))))
    t0 = time.perf_counter()
    for i in range(nb):
        b = (2 + i) % navail
        R["feed"](x0[b * v:(b + 1) * v])
    dt = time.perf_counter() - t0
    share = fft_share(O, W, isz, dt / nb)
    good = R["good"]()
    R["close"]()
    return {"value": nb * isz / dt / 1e6, "unit": "Msamples/s", "cores": R["cores"], "kind": R["kind"],
            "sample": "%d consecutive overlap-save blocks (%.1f Msamples) of the looped slab after 2 warm-up blocks; %s; forward FFT share of a block %s; %d CRC-good PDUs in the sample"
                      % (nb, nb * isz / 1e6, R["desc"], ("%.0f %%" % (100 * share)) if share is not None else "n/a", good),
            "seconds": dt, "host_cores": os.cpu_count() or 1}


def run_reference_arm(a, O, name, W, world):
    isz = O.geometry(W["sr"])[2].input_size
    P = plan_frames(O, W, isz)
    ncpu = os.cpu_count() or 1
    R = cpu_reference_setup(O, W, P)
    # size a step: one warm-up block timed, then ~1.5 s of CPU work per step, at most the slab
    nb_total_max = P["nblocks"]
    x = to_raw(O, W, render_range(O, W, P, 0, min(nb_total_max, 3) * isz, ncpu, W["seed"]))
    v = 2 * isz
    R["feed"](x[: v])
    t0 = time.perf_counter()
    R["feed"](x[v: 2 * v])
    per_block = time.perf_counter() - t0
    per = int(max(1, min(8, 1.5 / max(per_block, 1e-6))))
    need = (a.warmup + a.steps) * per
    have = min(need + 2, nb_total_max)
    if have * v > x.size:
        x = np.concatenate([x, to_raw(O, W, render_range(O, W, P, x.size // 2, have * isz - x.size // 2, ncpu, W["seed"] + 1))])
    pos = 2
    vals = []
    for s in range(a.warmup + a.steps):
        idx = [(pos + i) % have for i in range(per)]
        seg = np.concatenate([x[i * v:(i + 1) * v] for i in idx])
        t0 = time.perf_counter()
        R["feed"](seg)
        dt = time.perf_counter() - t0
        pos += per
        if s >= a.warmup:
            vals.append(dt)
    tot = sum(vals)
    val = a.steps * per * isz / tot / 1e6
    good = R["good"]()
    R["close"]()
    line = {"metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": config_of(name, W, P, world),
            "step": "%d consecutive overlap-save blocks (%.2f Msamples) of the slab per step (bounded sample of the workload)" % (per, per * isz / 1e6),
            "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": R["cores"], "kind": R["kind"], "sample": R["desc"]},
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "pdus_crc_good": good}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--sample-format", default=None, choices=sorted(SFMT))
    ap.add_argument("--multi", default="sharded", choices=["sharded", "broadcast"],
                    help="N > 1: 'sharded' = every GPU transforms 1/N of the blocks, spectrum slices all-to-all; 'broadcast' = every GPU transforms the whole capture")
    ap.add_argument("--loops", type=int, default=0, help="passes over the slab per step (0 = the workload's default)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: time the device-resident leg alone")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    name = a.workload or DEFAULT_BY_GPUS.get(max(world, a.gpus), "cfg3")
    W = prepare_workload(name, a, max(world, a.gpus))
    if a.warmup < 3:
        a.warmup = 3                                   # timing rule: at least three warm-up steps
    import orclib as O

    if a.impl == "reference":
        if rank == 0:
            run_reference_arm(a, O, name, W, max(world, a.gpus))
        return

    import torch
    import dumphfdl_b200 as hb
    hb.load()
    torch.cuda.set_device(local)
    numa = bind_near_gpu(local) if world > 1 else None     # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sharded = world > 1 and W["multi"] == "sharded"
    F = SFMT[W["sfmt"]]
    bps = F["bps"]
    ncpu = max(1, (os.cpu_count() or 1) // max(world, 1))
    g0 = O.geometry(W["sr"])[2]
    isz, ovl = g0.input_size, g0.overlap_length
    P = plan_frames(O, W, isz)
    nsamp, nblocks, loops = P["nsamp"], P["nblocks"], W["loops"]
    assert nsamp % world == 0
    part = nsamp // world
    # every rank renders its own 1/N of the slab (time slice); the whole slab is assembled over NCCL where needed
    x_part = to_raw(O, W, render_range(O, W, P, rank * part, part, ncpu, W["seed"] + 7919 * rank))       # flat I/Q values
    my_idx = list(range(rank, W["nch"], world))
    freqs = [P["freqs"][i] for i in my_idx]
    truth_set = set(t for t in P["truth"] if t[0] in set(freqs))
    ntruth = len(truth_set)
    h_part = torch.from_numpy(x_part.view(np.uint8)).pin_memory()      # raw bytes from here on (NCCL has no int16)
    d_part = torch.empty_like(h_part, device="cuda")
    d_part.copy_(h_part)
    nbuf = 3 if world > 1 else 1
    if sharded:
        # setup (untimed): the slab is assembled once so that every rank can cut out its share of each batch -- the blocks
        # [rank * bl, (rank + 1) * bl) plus the overlap in front of them (cyclic: the slab loops)
        bl = nblocks // world
        full = torch.empty(nsamp * bps, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(full, d_part)
        lo = bps * (rank * bl * isz - ovl)
        d_share = (torch.cat([full[lo:], full[: bps * bl * isz]]) if lo < 0 else full[lo: bps * (rank + 1) * bl * isz]).clone()
        del full
        h_share = torch.empty_like(d_share, device="cpu").pin_memory()
        h_share.copy_(d_share)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        bufs, d_slab = [], None
    else:
        bufs = [torch.empty(nsamp * bps, dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
        if world > 1:
            dist.all_gather_into_tensor(bufs[0], d_part)
            d_slab = bufs[0].clone() if rank == 0 else None
        else:
            bufs[0].copy_(d_part)
            d_slab = bufs[0]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def count(pdus):
        good = sum(1 for q in pdus if q.crc_good)
        exact = sum(1 for q in pdus if (q.freq, q.data()) in truth_set)
        return good, exact

    def make_frontend():
        fe_ = hb.Frontend(W["sr"], CF, freqs, sample_format=F["code"], device=local, max_blocks_per_batch=nblocks)
        if sharded:
            fe_.set_exchange(P["freqs"], world)
        return fe_

    class Exchange:
        """Sharded spectrum, host side of one rank: FFT of this rank's blocks -> all-to-all of the slices -> this rank's
        channels.  FFT + pack run on the main torch stream, the all-to-all and the hand-over to the demodulator on a side
        stream, so the FFT of pass i + 1 runs beside the exchange of pass i."""

        def __init__(self, fe_):
            self.fe = fe_
            n = world * bl * (W["nch"] // world) * int(fe_.L.hfdl_b200_slice_elems(fe_.h)) * 2
            self.send = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
            self.recv = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(nbuf)]
            self.fs = torch.cuda.Stream()            # FFT + pack (a real stream: handle 0 would mean "the frontend's own stream" to the C ABI)
            self.xs = torch.cuda.Stream()            # all-to-all + hand-over
            self.ev_fft = [torch.cuda.Event() for _ in range(2)]
            self.ev_sent = [torch.cuda.Event() for _ in range(2)]
            self.i = 0

        def run(self, d_in, ready=()):
            i = self.i
            for ev in ready:
                self.fs.wait_event(ev)                              # the samples of this pass have landed
            if i >= 2:
                self.fs.wait_event(self.ev_sent[i % 2])             # the all-to-all of pass i-2 has read this send buffer
            self.fe.spectrum_slices(d_in.data_ptr(), i * nblocks + rank * bl, bl, self.send[i % 2].data_ptr(), self.fs.cuda_stream)
            self.ev_fft[i % 2].record(self.fs)
            self.fe.wait_input(nbuf - 1)                            # the batch that read this receive buffer is through the channeliser
            with torch.cuda.stream(self.xs):
                self.xs.wait_event(self.ev_fft[i % 2])
                dist.all_to_all_single(self.recv[i % nbuf], self.send[i % 2])     # NCCL over NVLink, equal splits: part q -> rank q
                self.ev_sent[i % 2].record(self.xs)
                self.fe.process_slices(self.recv[i % nbuf].data_ptr(), nblocks, self.xs.cuda_stream)
            self.i += 1
            return self.ev_fft[i % 2]

    # ---- value: capture resident in HBM (sharded: every rank's share in its own HBM; broadcast: in rank 0's HBM)
    fe = make_frontend()
    state = {"pos": 0, "i": 0}
    xch = Exchange(fe) if sharded else None

    def pass_device():
        if sharded:
            xch.run(d_share)
            return
        if world > 1:
            buf = bufs[state["i"] % nbuf]
            fe.wait_input(nbuf - 1)                     # the batch that read this buffer nbuf passes ago is through the channeliser stage
            if rank == 0:
                buf.copy_(d_slab, non_blocking=True)
            dist.broadcast(buf, src=0)                  # NCCL over NVLink; only its stream is waited for, the batch pipeline keeps running
            torch.cuda.current_stream().synchronize()
        else:
            buf = d_slab
        fe.process_device(buf.data_ptr(), nsamp, state["pos"], nblocks)
        state["pos"] += nsamp
        state["i"] += 1

    for _ in range(a.warmup):
        for _ in range(loops):
            pass_device()
    torch.cuda.synchronize()
    fe.sync()
    fe.pdus()
    clk = ClockSampler(local)
    clk.start()
    l0 = fe.launches()
    barrier()
    fe.timer_start()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        for _ in range(loops):
            pass_device()
    torch.cuda.synchronize()
    ms_dev = fe.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    clocks = clk.stop()
    launches = fe.launches() - l0
    good, exact = count(fe.pdus())
    # per-kernel-class device time: a separate, untimed set of passes with an event pair around every launch (the event
    # records serialise the launch stream a little, so they stay out of the timed region)
    fe.profile(True)
    fe.profile_read()
    nprof = min(loops, 8)
    for _ in range(nprof):
        pass_device()
    torch.cuda.synchronize()
    prof = fe.profile_read()
    fe.profile(False)
    fe.pdus()
    ms = max(ms_dev, 0.0)
    if os.environ.get("HFDL_B200_DEBUG"):
        fe.L.hfdl_b200_print_summary(fe.h)
    geom = fe.geom
    fe.close()
    del xch

    # ---- e2e: from pinned host memory
    e2e_ms, e_good, d2h_per_step, h2d_per_step = float("nan"), 0, 0, 0
    if not a.skip_e2e:
        fe2 = make_frontend()
        st2 = {"pos": 0, "i": 0}
        xch2 = Exchange(fe2) if sharded else None

        copy_stream = torch.cuda.Stream() if world > 1 else None
        h_src = h_share if sharded else h_part
        d_parts = [torch.empty_like(h_src, device="cuda") for _ in range(2)] if world > 1 else None
        h2d_ev = [torch.cuda.Event(), torch.cuda.Event()] if world > 1 else None
        read_ev = [None, None]
        # sharded: a rank uploads exactly its own blocks; the overlap in front of them -- the last overlap_length samples of
        # the previous rank's blocks -- comes from that rank over NVLink (a ring shift per pass), so the capture crosses
        # PCIe exactly once in total.  Rank 0's overlap is the tail of the last rank's blocks of the PREVIOUS pass.
        ovb = ovl * bps
        shift = sharded and not os.environ.get("HFDL_BENCH_UPLOAD_OVERLAP")      # (A/B switch: every rank uploads its overlap itself)
        h_blocks = h_share[ovb:] if shift else None
        ovl_stream = torch.cuda.Stream() if sharded else None
        ovl_ev = [torch.cuda.Event(), torch.cuda.Event()] if sharded else None

        def start_h2d(i):
            # this rank's share of pass i over its own PCIe link, on its own stream (runs beside the FFT / all-gather of pass i-1)
            with torch.cuda.stream(copy_stream):
                if sharded and read_ev[i % 2] is not None:
                    copy_stream.wait_event(read_ev[i % 2])             # the FFT of pass i-2 has read this device buffer
                    if shift:
                        copy_stream.wait_event(ovl_ev[i % 2])          # ... and its tail has gone to the next rank
                if shift:
                    d_parts[i % 2][ovb:].copy_(h_blocks, non_blocking=True)
                else:
                    d_parts[i % 2].copy_(h_src, non_blocking=True)
                h2d_ev[i % 2].record(copy_stream)
            if shift:
                with torch.cuda.stream(ovl_stream):
                    ovl_stream.wait_event(h2d_ev[i % 2])
                    tgt = d_parts[i % 2] if rank > 0 else d_parts[(i + 1) % 2]
                    if rank == 0 and read_ev[(i + 1) % 2] is not None:
                        ovl_stream.wait_event(read_ev[(i + 1) % 2])    # the FFT that last read the other buffer (its overlap part is written now)
                    ops = [dist.P2POp(dist.isend, d_parts[i % 2][-ovb:], (rank + 1) % world),
                           dist.P2POp(dist.irecv, tgt[:ovb], (rank - 1) % world)]
                    for w in dist.batch_isend_irecv(ops):
                        w.wait()
                    ovl_ev[i % 2].record(ovl_stream)

        def pass_host():
            i = st2["i"]
            if sharded:
                if i == 0:
                    start_h2d(0)
                ready = [h2d_ev[i % 2]]
                if shift and rank > 0:
                    ready.append(ovl_ev[i % 2])
                elif shift and i > 0:
                    ready.append(ovl_ev[(i - 1) % 2])
                read_ev[i % 2] = xch2.run(d_parts[i % 2], ready)
                start_h2d(i + 1)
            elif world > 1:
                buf = bufs[i % nbuf]
                if i == 0:
                    start_h2d(0)
                fe2.wait_input(nbuf - 1)
                torch.cuda.current_stream().wait_event(h2d_ev[i % 2])
                dist.all_gather_into_tensor(buf, d_parts[i % 2])       # NVLink: every GPU ends up with the whole capture
                torch.cuda.current_stream().synchronize()
                start_h2d(i + 1)                                       # d_parts[(i+1) % 2] was read by the all-gather of pass i-1, which has completed
                fe2.process_device(buf.data_ptr(), nsamp, st2["pos"], nblocks)
            else:
                # H2D inside the C-ABI call; the slab is never rewritten, so the call need not wait for its last copy (the PDU
                # pick-up below then runs beside it); the final flush waits for everything
                fe2.push_ptr(h_part.data_ptr(), nsamp, wait=False)
            st2["pos"] += nsamp
            st2["i"] += 1

        for _ in range(max(1, a.warmup * loops // 4)):
            pass_host()
        torch.cuda.synchronize()
        fe2.flush()
        fe2.pdus()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            for _ in range(loops):
                pass_host()
                e_good += count(fe2.pdus())[0]           # PDU records of the batches that finished meanwhile (D2H)
        torch.cuda.synchronize()
        fe2.flush()                                      # the timed region ends when every PDU is on the host
        e_good += count(fe2.pdus())[0]
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        d2h_per_step = fe2.result_bytes_per_batch() * loops
        h2d_per_step = (h_blocks.numel() if shift else h_src.numel()) * loops
        fe2.close()
        del xch2

    t = torch.tensor([ms, e2e_ms if e2e_ms == e2e_ms else 0.0, float(good), float(exact), float(e_good), float(ntruth), float(launches),
                      float(h2d_per_step), float(d2h_per_step)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(tmax[0]), float(tmax[1])
        good, exact, e_good, ntruth, launches = float(tsum[2]), float(tsum[3]), float(tsum[4]), float(tsum[5]), float(tsum[6])
        h2d_per_step, d2h_per_step = float(tsum[7]), float(tsum[8])
    else:
        ms, e2e_ms, ntruth = float(t[0]), float(t[1]), float(t[5])
    total_samples = a.steps * loops * nsamp              # slab samples: every GPU sees the whole capture, it is counted once
    value = total_samples / (ms / 1e3) / 1e6
    e2e = total_samples / (e2e_ms / 1e3) / 1e6 if e2e_ms > 0 else None
    if rank == 0:
        N, M, out = geom.fft_size, geom.fft_inv_size, geom.out_per_block
        Cn = len(freqs)
        # algorithmic (compulsory) HBM bytes per overlap-save block ON ONE GPU, SURVEY 8(d): ingest read + spectrum write
        # (this GPU's share of the blocks: all of them unless the spectrum is sharded) + this GPU's channels' spectrum
        # slice read + tap slice read + baseband write + demod read
        fsh = 1.0 / world if sharded else 1.0
        b_blk = int(fsh * (isz * bps + N * 8)) + Cn * M * 8 + Cn * M * 8 + Cn * out * 16
        nout = out * geom.resamp_rate
        alg = {"fft_pass1": isz * bps * fsh, "fft_pass2": N * 8 * fsh if geom.fft_passes == 2 else 0, "fft_pass3": N * 8 * fsh if geom.fft_passes == 3 else 0,
               "chan_extract": Cn * M * 16 + Cn * out * 8, "resamp": Cn * out * 8 * (1 + geom.resamp_rate),
               "agc": Cn * nout * (8 + 12), "bank": Cn * nout * (8 + 8 + 256), "loop": Cn * nout * (256 + 4), "fec": 0}
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        kern = {k: {"ms_total": v[0], "launches": v[1], "ms_per_launch": (v[0] / v[1] if v[1] else 0.0)} for k, v in prof.items()}
        traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                tj = json.load(f)
            if tj.get("workload") == "%s-%s" % (name, W["sfmt"]) and tj.get("blocks_per_batch") == nblocks and tj.get("channels") == Cn and world == 1:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        nbatches = nprof
        roofs = {}
        for k, v in prof.items():
            if not v[1] or not alg.get(k):
                continue
            per_launch_bytes = alg[k] * nblocks * nbatches / v[1]
            dur = v[0] / v[1] / 1e3
            ach = per_launch_bytes / dur / 1e9
            roofs[k] = {"achieved": round(ach, 1), "frac": round(ach / peak, 4), "traffic": traffic.get(k), "algorithmic_bytes_per_launch": int(per_launch_bytes),
                        "avg_launch_ms": round(dur * 1e3, 4), "launches_per_batch": v[1] / nbatches}
        # the dominant kernel and the contract's roofline for it: the WHOLE path's algorithmic bytes of the blocks one launch
        # covers (B_blk x blocks per launch) over that kernel's average launch duration
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0] if prof else None
        roof = None
        if dom and prof[dom][1]:
            blocks_per_launch = nblocks * nbatches / prof[dom][1]
            dur = prof[dom][0] / prof[dom][1] / 1e3
            ach = b_blk * blocks_per_launch / dur / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic.get(dom),
                    "peak_source": peak_src, "algorithmic_bytes_per_block": b_blk, "blocks_per_launch": blocks_per_launch, "avg_launch_ms": dur * 1e3,
                    "note": ("the step costs what its slowest pipeline stage costs; %s is that stage. It is a latency-bound sequential recurrence (one warp chain per "
                             "channel), so this fraction says how far the whole path is from the HBM roofline, not how busy HBM is; the kernels' own bytes are under "
                             "roofline_kernels" % dom) if dom in ("loop", "agc", "fec") else None}
        pipe_ach = b_blk * nblocks * nbatches / (ms / 1e3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_of(name, W, P, world),
                "step": "%d passes over the slab (%.1f Msamples, %d batches of %d blocks per GPU)" % (loops, loops * nsamp / 1e6, loops, nblocks),
                "scaling_note": "the workload is BASELINE's configuration for the GPU count (cfg3 at 1-2, cfg4 at 4, cfg5 at 8 GPUs): capture rate and channel count grow with N; "
                                "value counts capture samples once" + ("" if sharded or world == 1 else " although every GPU transforms the whole capture"),
                "pdus_per_s": good / (ms / 1e3), "pdus_crc_good": good, "pdus_exact": exact, "pdus_expected_per_pass": ntruth,
                "e2e": {"value": e2e, "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": int(d2h_per_step),
                        "pdus_per_s": (e_good / (e2e_ms / 1e3)) if e2e_ms > 0 else None, "ms_per_step": e2e_ms / a.steps if e2e_ms > 0 else None},
                "gpu_launches": int(launches), "clocks": clocks, "kernels": kern, "roofline": roof, "roofline_kernels": roofs,
                "roofline_pipeline": {"bound": "hbm", "achieved": pipe_ach, "peak": peak, "unit": "GB/s", "frac": pipe_ach / peak, "algorithmic_bytes_per_block": b_blk},
                "wall_ms_per_step": 1e3 * wall / a.steps}
        if numa:
            line["host_binding"] = numa
        if not a.no_cpu_baseline and world == 1:          # the CPU baseline is reported by the single-GPU run only
            line["cpu_baseline"] = cpu_reference(O, W, P, isz, x_part)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
