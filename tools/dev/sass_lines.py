"""Dev-time: static SASS instruction count per source line of one kernel (nvdisasm -g -c output).
  nvcc ... -lineinfo -cubin -o k.cubin recon_kernels.cu && nvdisasm -g -c k.cubin > k.sass
  python tools/dev/sass_lines.py k.sass recon_kernel3ILi1 [file-filter]"""
import collections
import re
import sys

path, kernel = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else None
lines = open(path).read().split("\n")
on, cur = False, None
hist, ops = collections.Counter(), collections.Counter()
for l in lines:
    if l.startswith(".text."):
        on = kernel in l
        continue
    if not on:
        continue
    m = re.search(r'//## File ".*?([\w\.]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_\.]+)", l)
    if m and cur:
        hist[cur] += 1
        ops[m.group(2).split(".")[0]] += 1
print("total static instructions", sum(hist.values()))
for k, v in sorted(hist.items()):
    if flt is None or flt in k[0]:
        print("%-28s %5d  %d" % (k[0], k[1], v))
print(ops.most_common(50))
