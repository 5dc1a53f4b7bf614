"""Writes profiles/r01_ncu_summary.md and profiles/r01_traffic.json from `ncu -i <rep> --page raw --csv` exports
(usage: python profiles/ncu_summary.py loop_raw.csv rest_raw.csv, run inside the directory holding the exports)."""
import csv, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
rows = []
for f in sys.argv[1:]:
    r = list(csv.reader(open(f))); h = r[0]
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
            'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum']
    idx = {k: h.index(k) for k in keys if k in h}
    units = {k: r[1][i] for k, i in idx.items()}
    for row in r[2:]:
        rows.append({k: row[i] for k, i in idx.items()} | {'_units': units})
tobytes = lambda v, u: float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
toms = lambda v, u: float(v) * {'us': 1e-3, 'ms': 1, 'ns': 1e-6, 's': 1e3}[u]
seen = {}
lines = ["# ncu --set full --clock-control none captures, round 1 (B200, cfg2 bench: 2 Msps CF32, 8 channels, 88 blocks per step)", "",
         "Source reports: `gpurun_out/r01_loop.ncu-rep`, `gpurun_out/r01_rest.ncu-rep` (scratch, not tracked); command lines in profiles/README.md.", "",
         "| kernel | grid x block | regs | duration ms | dram read MB | dram write MB | dram % of peak | sm throughput % | warps active % | warp instr |",
         "|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
name_map = {'loop_kernel': 'loop', 'fft_col_pass': 'fft_pass1', 'fft_row_pass': 'fft_pass2', 'fft_col_pass_reg': 'fft_pass1', 'fft_row_pass_reg': 'fft_pass2', 'fft_last_pass_nat': 'fft_pass2', 'chan_extract': 'chan_extract', 'agc_kernel': 'agc',
            'bank_kernel': 'bank', 'fec_kernel': 'fec', 'resamp_kernel': 'resamp'}
for d in rows:
    kn = d['Kernel Name'].split('(')[0].replace('void ', '').strip()
    if kn in seen:
        continue
    seen[kn] = 1
    u = d['_units']
    rd = tobytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']); wr = tobytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
    ms = toms(d['gpu__time_duration.sum'], u['gpu__time_duration.sum'])
    lines.append("| %s | %s x %s | %s | %.4f | %.2f | %.2f | %.2f | %.1f | %.1f | %d |" % (
        kn, d['launch__grid_size'], d['launch__block_size'], d['launch__registers_per_thread'], ms, rd / 1e6, wr / 1e6,
        float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']), float(d['sm__throughput.avg.pct_of_peak_sustained_elapsed']),
        float(d['sm__warps_active.avg.pct_of_peak_sustained_active']), float(d.get('smsp__inst_executed.sum', 0))))
    traffic[name_map.get(kn.split('<')[0], kn)] = int(rd + wr)
open(os.path.join(HERE, 'r01_ncu_summary.md'), 'w').write("\n".join(lines) + "\n")
json.dump({"workload": "cfg2", "blocks_per_step": 88, "channels": 8,
           "source": "ncu --set full --clock-control none, one launch per kernel class inside `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` "
                     "(profiles/r01_ncu_summary.md); dram__bytes_read.sum + dram__bytes_write.sum",
           "dram_bytes_per_launch": traffic}, open(os.path.join(HERE, 'r01_traffic.json'), 'w'), indent=1)
print("\n".join(lines[5:]))
