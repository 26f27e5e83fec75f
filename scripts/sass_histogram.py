"""SASS evidence for profiles/: per kernel of libvlite_b200.so, how often the mnemonics that prove the
Blackwell-native paths (and the integer hot loop) occur.  No GPU needed (cuobjdump on the built library).
usage: python scripts/sass_histogram.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vlite_b200", "lib", "libvlite_b200.so")
COLS = ["UTCIMMA", "UTCOMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UTMALDG", "UBLKCP", "SYNCS", "UCGABAR", "PRMT", "LOP3", "POPC",
        "IDP", "IMAD", "VIMNMX3", "FMNMX3|FMNMX", "STTM", "DFMA", "STS", "LDS", "REDG|RED|ATOM"]
NOTE = {"UTCIMMA": "tcgen05.mma kind::i8", "UTCOMMA": "tcgen05.mma kind::mxf4 block_scale", "UTCBAR": "tcgen05.commit", "LDTM": "tcgen05.ld", "UTCATOMSWS": "tcgen05.alloc",
        "UTMALDG": "cp.async.bulk.tensor (TMA)", "UBLKCP": "cp.async.bulk (TMA 1-D)", "SYNCS": "mbarrier ops",
        "UCGABAR": "barrier.cluster", "IDP": "dp4a", "VIMNMX3": "3-input max", "DFMA": "fp64 fma"}

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names, counts, sizes = [], {}, {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        names.append(cur)
        counts[cur] = collections.Counter()
        sizes[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        sizes[cur] += 1
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
short = {}
for n, d in zip(names, dem):
    d = re.sub(r"^void ", "", d)
    d = re.sub(r"\(CUtensorMap_st.*|\(vl::.*|\((?:unsigned|const|int|float|long|uint|void).*", "", d)
    short[n] = d.replace("(int)", "").replace("(bool)", "")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_histogram.md")
with open(out, "w") as f:
    f.write("# SASS mnemonic counts per kernel of libvlite_b200.so (cuobjdump -sass, sm_100a)\n\n")
    f.write("Static instruction counts (not executed counts).  " + "; ".join(f"`{k}` = {v}" for k, v in NOTE.items()) + ".\n")
    f.write("No `HMMA`/`IMMA`/`HGMMA` anywhere: the only tensor-core path is tcgen05 (`UTCIMMA` int8, `UTCOMMA` block-scaled e2m1).\n\n")
    f.write("| kernel | SASS instrs | " + " | ".join(c.split("|")[0] for c in COLS) + " |\n|---|---:|" + "---:|" * len(COLS) + "\n")
    legacy = 0
    for n in sorted(names, key=lambda x: short[x]):
        c = counts[n]
        legacy += sum(v for k, v in c.items() if k in ("HMMA", "IMMA", "HGMMA", "IGMMA", "QGMMA"))
        row = []
        for col in COLS:
            row.append(sum(v for k, v in c.items() if re.fullmatch(col, k) or k.startswith(col.split("|")[0]) and "|" not in col
                           or ("|" in col and re.fullmatch(col, k))))
        f.write(f"| `{short[n]}` | {sizes[n]} | " + " | ".join(str(x) if x else "" for x in row) + " |\n")
    f.write(f"\nlegacy tensor instructions (HMMA/IMMA/*GMMA) in the library: {legacy}\n")
print(out)
