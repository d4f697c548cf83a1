"""Per-kernel counts of the SASS mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), UTMALDG (TMA tensor
loads), LDGSTS (cp.async), SYNCS (mbarrier), RED / ATOM, LDG.E.128, MUFU.

    python tools/sass_counts.py [sml_b200/libsml_b200.so] > profiles/r02_sass.md
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "sml_b200/libsml_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
pats = [("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("UBLKCP", r"\bUBLKCP"),
        ("UTMALDG", r"\bUTMALDG"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("RED", r"\bRED\."), ("ATOM", r"\bATOM[GS]?\."),
        ("LDG.128", r"\bLDG\.E\.(?:\w+\.)*128"), ("STG.128", r"\bSTG\.E\.(?:\w+\.)*128"), ("MUFU", r"\bMUFU"), ("FFMA", r"\bFFMA"), ("total", r"^\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]")]
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*$", "", name).replace("void ", "")
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for k, p in pats:
        if re.search(p, line):
            counts[name][k] += 1
print("# SASS mnemonic counts per kernel (`cuobjdump -sass %s`, sm_100a)\n" % so)
print("| kernel | " + " | ".join(k for k, _ in pats) + " |")
print("|---|" + "---:|" * len(pats))
tot = collections.Counter()
for n, c in sorted(counts.items()):
    print("| `%s` | " % n + " | ".join(str(c[k]) if c[k] else "" for k, _ in pats) + " |")
    tot.update(c)
print("| **all kernels** | " + " | ".join(str(tot[k]) for k, _ in pats) + " |")
