"""Dev-time: program-order SASS listing of one kernel with the source line of every instruction.
  python tools/dev/sass_listing.py k.sass recon_kernel3ILi1 > listing.txt"""
import re
import sys

path, kernel = sys.argv[1], sys.argv[2]
on, cur = False, ""
for l in open(path).read().split("\n"):
    if l.startswith(".text."):
        on = kernel in l
        continue
    if not on:
        continue
    m = re.search(r'//## File ".*?([\w\.]+)", line (\d+)(.*)', l)
    if m:
        f = m.group(1).replace("recon_kernels.cu", "k")
        inl = re.findall(r'inlined at ".*?([\w\.]+)", line (\d+)', m.group(3))
        cur = "%s:%s" % (f, m.group(2)) + "".join(" <%s:%s" % (a.replace("recon_kernels.cu", "k"), b) for a, b in inl)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        print("%s  %-70s  %s" % (m.group(1), m.group(2).strip(), cur))
    elif re.match(r"^\.L_x_\d+:", l):
        print(l.strip())
