"""Per-kernel SASS mnemonic counts of libmiphei_b200.so (cuobjdump -sass): the instructions that prove tcgen05 / TMEM / TMA.
   python tools/sass_counts.py > profiles/rNN_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "miphei-vit_b200", "libmiphei_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "REDG", "HMMA", "MUFU", "SYNCS"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for k in KEYS:
        if op == k or op.startswith(k + "."):
            counts[cur][k] += 1
    if op.startswith("UTCHMMA") and ".2CTA" in op:
        counts[cur]["UTCHMMA.2CTA"] += 1
print("# SASS mnemonic counts per kernel of %s (cuobjdump -sass; sm_100a)" % os.path.basename(lib))
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor load / store, REDG = global reduction")
print("%-110s %s" % ("kernel", " ".join("%12s" % k for k in KEYS)))
tot = collections.Counter()
for fn, c in counts.items():
    name = demangle(fn)
    name = re.sub(r"\(.*$", "", name).replace("mv::", "")
    if not any(c[k] for k in KEYS):
        continue
    print("%-110s %s" % (name[:110], " ".join("%12d" % c[k] for k in KEYS)))
    tot.update(c)
print("%-110s %s" % ("TOTAL", " ".join("%12d" % tot[k] for k in KEYS)))
