"""cuobjdump -sass libnbpc.so | python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt
Counts the SASS mnemonics that prove the Blackwell-native path per kernel (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, mbarrier -> SYNCS; HMMA (legacy mma.sync) must not appear."""
import collections
import re
import subprocess
import sys

cur, cnt = None, collections.OrderedDict()
MN = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for mn in MN:
        if re.search(r"\b" + mn + r"\b", line):
            cnt[cur][mn] += 1
print("SASS mnemonics per kernel (cuobjdump -sass libnbpc.so): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG,")
print("mbarrier -> SYNCS, cp.async -> LDGSTS; kernels without tensor-core / TMA instructions are omitted.")
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
hmma = 0
for (k, c), name in zip(cnt.items(), names):
    hmma += c.get("HMMA", 0)
    if any(c.get(x) for x in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG")):
        print(f"{name.split('(')[0][:78]:78s} " + ", ".join(f"{m} x{c[m]}" for m in MN if c.get(m)))
print(f"total HMMA (legacy mma.sync) instructions in the library: {hmma}")
