"""Dev-time: join an `ncu --page source --csv` export (SASS view, per-instruction counters) with the nvdisasm
listing of the same kernel (tools/dev/sass_listing.py) and sum executed warp instructions / stall samples / shared
wavefronts per source line.   python tools/dev/ncu_by_line.py src.csv listing.txt [n_macroblocks]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iex, isamp, iwave = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
ins = [r for r in rows[2:] if len(r) > iex]
lst = [l for l in open(sys.argv[2]).read().split("\n") if l and not l.startswith(".L_x")]
assert len(ins) == len(lst), (len(ins), len(lst))
nmb = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
agg = collections.OrderedDict()
for r, l in zip(ins, lst):
    src = l.split("  ")[-1].strip()
    key = src.split(" <")[-1] if " <" in src else src      # attribute inlined helpers to the outermost call site (the kernel body)
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += int(r[iex]); a[1] += int(r[isamp]); a[2] += int(r[iwave] or 0)
tot = [sum(a[i] for a in agg.values()) for i in range(3)]
print("total: %.1f instr/MB, %d samples, %.1f wavefronts/MB" % (tot[0] / nmb, tot[1], tot[2] / nmb))
def keyf(k):
    f, _, n = k.partition(":")
    return (f, int(n) if n.isdigit() else 0)
for k in sorted(agg, key=keyf):
    a = agg[k]
    if a[0]:
        print("%-28s %8.2f instr/MB  %5.1f%% samples  %6.2f wavefronts/MB" % (k, a[0] / nmb, 100.0 * a[1] / max(tot[1], 1), a[2] / nmb))
