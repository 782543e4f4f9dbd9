#!/usr/bin/env python
"""Counts of the Blackwell-specific SASS mnemonics per kernel of liboc_b200.so (cuobjdump -sass): the evidence that the kernels
use the bulk-copy engine (UBLKCP), cp.async (LDGSTS), mbarriers (SYNCS), tcgen05 MMAs (UTCHMMA) and tensor memory (LDTM / STTM).
usage: sass_markers.py [lib.so] > profiles/rNN_sass_markers.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "jaxovercooked_b200", "liboc_b200.so")
MARKS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "FHFMA", "HMMA", "SHFL", "VOTE",
         "ACQBULK", "CCTL", "ERRBAR", "MEMBAR", "FENCE", "ATOMG", "RED", "STL", "LDL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        counts[cur] = collections.Counter(); total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for k in MARKS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# {os.path.basename(lib)}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a); STL/LDL = local-memory spills")
for k, c in counts.items():
    marks = "  ".join(f"{m}={c[m]}" for m in MARKS if c[m])
    print(f"{k:60s} instrs={total[k]:6d}  {marks}")
