"""Join an ncu `--page source --csv --print-source sass` export with nvdisasm line info of the same cubin and print the
executed warp instructions per source line / per line range:
  cuobjdump -xelf all build/knn.o; nvdisasm --print-line-info knn.sm_100a.cubin > all.sass
  python tools/sass_lines.py all.sass profile.csv <mangled kernel name prefix> <source file> [queries] [a-b ...]"""
import collections
import csv
import re
import sys

sass, prof_csv, kern, srcfile = sys.argv[1:5]
per = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
ranges = [tuple(int(v) for v in a.split("-")) for a in sys.argv[6:]]
lines = open(sass).read().split("\n")
i0 = [i for i, l in enumerate(lines) if ".section\t.text." + kern in l][0]
cur, seq = None, []
base = srcfile.split("/")[-1]
for l in lines[i0 + 1:]:
    if l.lstrip().startswith(".section"):
        break
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
    if m:
        if m.group(1) == base:
            cur = int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
        seq.append(cur)
rows = list(csv.reader(open(prof_csv)))
ia = rows[1].index("Instructions Executed")
prof = [int(r[ia]) for r in rows[2:] if r[ia].isdigit()]
assert len(prof) == len(seq), (len(prof), len(seq))
by = collections.Counter()
for ln, n in zip(seq, prof):
    by[ln] += n
src = open(srcfile).read().split("\n")
print("total", sum(by.values()) / per)
if ranges:
    for a, b in ranges:
        print(f"{a}-{b}: {sum(n for ln, n in by.items() if ln and a <= ln <= b) / per:8.1f}")
else:
    for ln, n in sorted(by.items(), key=lambda x: -x[1])[:50]:
        print(f"{n / per:8.1f}  {ln}: {src[ln - 1].strip()[:100] if ln else ''}")
