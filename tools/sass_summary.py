"""Developer tool: per-kernel counts of the Blackwell-specific SASS mnemonics in the in-tree library
(cuobjdump -sass): UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (tcgen05.ld / st),
UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc), SYNCS (mbarrier), plus HMMA (legacy mma.sync -- expected 0).
Usage: python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
    "super_resolution-image-reconstructer-multi_generator_gan_b200", "libsrgan_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "ACQBULK",
      "UBLKCP", "UTMACMDFLUSH", "FENCE", "ELECT", "NANOSLEEP"]
counts = collections.OrderedDict()
cur = None
arch = None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        cur = name.split("(")[0][:70]
        counts[cur] = collections.Counter()
        counts[cur]["_insts"] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_insts"] += 1
        for k in MN:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# SASS summary of {os.path.basename(so)} (arch {arch}; cuobjdump -sass; one row per kernel with any tensor / TMA / TMEM instruction)")
print("# tcgen05.mma = UTCHMMA, TMA load / store = UTMALDG / UTMASTG, tcgen05.ld / st = LDTM / STTM, tcgen05.commit = UTCBAR,")
print("# TMEM alloc = UTCATOMSWS, mbarrier = SYNCS; HMMA = legacy mma.sync (none expected)")
cols = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA"]
print("| kernel | SASS insts | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
tot = collections.Counter()
n_k = 0
for k, c in counts.items():
    n_k += 1
    for x in cols:
        tot[x] += c[x]
    if any(c[x] for x in cols[:7]):
        print(f"| `{k}` | {c['_insts']} | " + " | ".join(str(c[x]) for x in cols) + " |")
print(f"| **all {n_k} kernels** | {sum(c['_insts'] for c in counts.values())} | " + " | ".join(str(tot[x]) for x in cols) + " |")
