"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: stall samples and executed instructions
per CUDA source line (rows whose Address column is '-' are the per-line aggregates)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = None; cur = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        out.append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), cur, d["Line No"], r[1].strip(), dict(zip(hdr[4:], r[4:]))))
tot = sum(o[0] for o in out); ti = sum(o[1] for o in out)
print("total samples %d, warp instructions %d" % (tot, ti))
for s, i, f, ln, src, d in sorted(out, key=lambda x: -x[0])[:top]:
    st = sorted(((int(v or 0), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:3]
    print("%5d %5.1f%% inst %7d  %s:%s  %s   [%s]" % (s, 100.0 * s / max(tot, 1), i, f, ln, src[:110], ", ".join("%s %d" % (k[6:], v) for v, k in st if v)))
