#!/usr/bin/env python
"""SASS opcode evidence per translation unit (no GPU needed): which kernels of libdmvae_b200 contain the Blackwell-native
instructions (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor copies, UBLKCP =
bulk copies, SYNCS = mbarrier) and the legacy warp-level HMMA of the ELBO contraction.
    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "deep-mixture-vae_b200", "build")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCATOMSWS", "SYNCS", "HMMA",
        "LDGSTS", "MUFU", "REDG", "ATOMG"]


def main():
    for obj in sorted(os.listdir(OBJ)):
        if not obj.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        per = collections.OrderedDict()
        cur = None
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                per[cur] = collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur:
                per[cur][m.group(2)] += 1
        print("==== %s: %d kernels ====" % (obj, len(per)))
        tot = collections.Counter()
        for k, c in per.items():
            tot.update(c)
        print("  all kernels: " + ", ".join("%s x%d" % (k, tot[k]) for k in KEYS if tot[k]))
        for k, c in per.items():
            hits = ", ".join("%s x%d" % (kk, c[kk]) for kk in KEYS if c[kk])
            if any(c[kk] for kk in KEYS[:10]) or c["HMMA"]:
                name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0][:90]
                print("    %-92s %s" % (name, hits))


if __name__ == "__main__":
    main()
