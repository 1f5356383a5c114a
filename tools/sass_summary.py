#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit),
UCGABAR (cluster barrier), HMMA (mma.sync), SYNCS (mbarrier).   python tools/sass_summary.py [lib.so] > profiles/sass_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rag_gesture_b200", "librg_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2D.2CTA", "UTMASTG", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "UCGABAR",
        "HMMA", "SYNCS", "LDGSTS", "UBLKCP"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = name.replace("void (anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or op.startswith(k + "_") or (k.count(".") and op.startswith(k)):
                counts[name][k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass), kernels with none of them omitted")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UCGABAR", "HMMA", "SYNCS"]))
tot = collections.Counter()
for n, c in counts.items():
    if not any(c.values()):
        continue
    tot.update(c)
    tag = " [2CTA]" if c["UTCHMMA.2CTA"] else ""
    print(f"{(n + tag)[:70]:70s} " + " ".join(f"{c[k]:8d}" for k in ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UCGABAR", "HMMA", "SYNCS"]))
print(f"{'TOTAL':70s} " + " ".join(f"{tot[k]:8d}" for k in ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UCGABAR", "HMMA", "SYNCS"]))
print(f"# of which cta_group::2: UTCHMMA.2CTA {tot['UTCHMMA.2CTA']}, UTMALDG.2D.2CTA {tot['UTMALDG.2D.2CTA']}, UTCBAR.2CTA.MULTICAST {tot['UTCBAR.2CTA.MULTICAST']}")
