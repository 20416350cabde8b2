"""SASS opcode census of libbbbp_b200.so: which kernels carry tcgen05 (UTC*MMA), TMEM loads / stores (LDTM / STTM), TMA
(UTMALDG / UTMASTG / UBLKCP), legacy warp MMA (HMMA), cp.async (LDGSTS).  Writes profiles/r02_sass_census.txt.

    python tools/sass_census.py
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bbbp-multi-modal-deep-ensemble-framework_b200", "libbbbp_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "LDGSTS", "F2FP.SATFINITE",
       "SYNCS", "ELECT"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("bbbp::", "")
        counts[cur] = collections.Counter()
        continue
    if cur:
        for op in OPS:
            if re.search(r"\b" + re.escape(op) + r"\b|\b" + re.escape(op) + r"\.", line):
                counts[cur][op] += 1
rows = [(k, c) for k, c in counts.items() if any(c[o] for o in OPS if o not in ("SYNCS", "ELECT", "F2FP.SATFINITE"))]
out = [f"# python tools/sass_census.py -- cuobjdump -sass {os.path.basename(LIB)}: opcode counts per kernel ({len(counts)} kernels in the library,",
       f"# {len(rows)} with tensor-core / TMEM / TMA / cp.async instructions listed)",
       f"{'kernel':78s} " + " ".join(f"{o:>8s}" for o in OPS)]
for k, c in sorted(rows, key=lambda kc: (-kc[1]['UTCHMMA'], kc[0])):
    out.append(f"{k[:78]:78s} " + " ".join(f"{c[o]:8d}" for o in OPS))
text = "\n".join(out) + "\n"
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", "r02_sass_census.txt"), "w").write(text)
print(text)
