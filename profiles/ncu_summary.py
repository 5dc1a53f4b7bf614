"""Writes profiles/<tag>_ncu_summary.md and profiles/<tag>_traffic.json from `ncu -i <rep> --page raw --csv` exports.
usage: python profiles/ncu_summary.py <tag> <workload> <blocks_per_batch> <channels> <raw.csv>...   (first occurrence of a kernel
class wins, except FFT passes / chan_extract, whose launches of one batch are summed: they run per sub-batch)"""
import csv, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
tag, workload, nblocks, nch = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = []
for f in sys.argv[5:]:
    r = list(csv.reader(open(f)))
    h = r[0]
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
            'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct']
    idx = {k: h.index(k) for k in keys if k in h}
    units = {k: r[1][i] for k, i in idx.items()}
    for row in r[2:]:
        rows.append({k: row[i] for k, i in idx.items()} | {'_units': units})
tobytes = lambda v, u: float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
toms = lambda v, u: float(v) * {'us': 1e-3, 'ms': 1, 'ns': 1e-6, 's': 1e3, 'usecond': 1e-3, 'msecond': 1, 'nsecond': 1e-6, 'second': 1e3}[u]
name_map = {'loop_kernel': 'loop', 'fft_col_pass_reg': 'fft_pass', 'fft_last_pass_nat': 'fft_last_pass', 'chan_extract': 'chan_extract', 'agc_kernel': 'agc',
            'bank_kernel': 'bank', 'fec_kernel': 'fec', 'resamp_kernel': 'resamp'}
lines = ["# ncu --set full --clock-control none captures, %s (B200, %s bench: %d blocks per batch, %d channels)" % (tag, workload, nblocks, nch), "",
         "One line per captured launch (kernels run serialised and cold under ncu: compare shares and DRAM bytes, not absolute times).", "",
         "| kernel | grid x block | regs | duration ms | dram read MB | dram write MB | dram % of peak | L2 hit % | sm throughput % | warps active % | warp instr |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
traffic, count = {}, {}
col_seen = 0             # column passes seen since the last "last pass": the first is pass 1, the second pass 2
for d in rows:
    kn = d['Kernel Name'].split('(')[0].replace('void ', '').strip()
    u = d['_units']
    rd = tobytes(d['dram__bytes_read.sum'], u['dram__bytes_read.sum']); wr = tobytes(d['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
    ms = toms(d['gpu__time_duration.sum'], u['gpu__time_duration.sum'])
    lines.append("| %s | %s x %s | %s | %.4f | %.2f | %.2f | %.2f | %s | %.1f | %.1f | %d |" % (
        kn, d['launch__grid_size'], d['launch__block_size'], d['launch__registers_per_thread'], ms, rd / 1e6, wr / 1e6,
        float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']), d.get('lts__t_sector_hit_rate.pct', '-'), float(d['sm__throughput.avg.pct_of_peak_sustained_elapsed']),
        float(d['sm__warps_active.avg.pct_of_peak_sustained_active']), float(d.get('smsp__inst_executed.sum', 0))))
    cls = name_map.get(kn.split('<')[0], kn)
    if cls == 'fft_pass':
        col_seen += 1
        cls = 'fft_pass%d' % col_seen
    elif cls == 'fft_last_pass':
        cls = 'fft_pass%d' % (col_seen + 1)
        col_seen = 0
    traffic[cls] = traffic.get(cls, 0) + int(rd + wr)
    count[cls] = count.get(cls, 0) + 1
open(os.path.join(HERE, tag + '_ncu_summary.md'), 'w').write("\n".join(lines) + "\n")
json.dump({"workload": workload, "blocks_per_batch": nblocks, "channels": nch,
           "source": "ncu --set full --clock-control none, the launches of one batch inside `python bench.py --steps 1 --warmup 3 --loops 1 --no-cpu-baseline --skip-e2e` "
                     "(profiles/%s_ncu_summary.md); dram__bytes_read.sum + dram__bytes_write.sum, summed over the launches of a class within the batch" % tag,
           "launches_in_capture": count, "dram_bytes_per_batch": traffic,
           "dram_bytes_per_launch": {k: int(v / count[k]) for k, v in traffic.items()}}, open(os.path.join(HERE, tag + '_traffic.json'), 'w'), indent=1)
print("\n".join(lines[5:]))
