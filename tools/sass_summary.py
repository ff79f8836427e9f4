"""Per-kernel SASS evidence of libgripb200.so (run here, no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts, per kernel, the mnemonics that prove the Blackwell path (B200_PROFILING.md): UTCHMMA (tcgen05.mma kind::f16,
.2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA load / store), UTCBAR (tcgen05.commit), SYNCS
(mbarrier), HMMA (legacy mma.sync), MUFU, HFMA2 (packed fp16 math)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "menghini-neurips23-code_b200", "libgripb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDSM", "MUFU", "HFMA2",
        "HMUL2", "FFMA", "ACQBULK", "UTMACMDFLUSH", "USETMAXREG", "total"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("gb::", "")
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        if op.startswith("UTCHMMA"):
            cur["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
            continue
        for k in keys:
            if k not in ("UTCHMMA.2CTA", "UTCHMMA", "total") and op.startswith(k):
                cur[k] += 1
                break
print(f"# cuobjdump -sass {os.path.basename(lib)} — instruction counts per kernel (static); arch sm_100a")
print("# " + " ".join(f"{k:>9s}" if i else f"{k:>12s}" for i, k in enumerate(keys)) + "  kernel")
for name, c in per.items():
    print("  " + " ".join(f"{c[k]:9d}" if i else f"{c[k]:12d}" for i, k in enumerate(keys)) + "  " + name)
