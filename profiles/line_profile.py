#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples: joins an .ncu-rep SASS page with `nvdisasm -g` line info.
usage: line_profile.py rep.ncu-rep lib.so kernel_substr [ntop]"""
import csv, io, re, subprocess, sys, tempfile, os, glob, collections
rep, lib, ksub = sys.argv[1:4]; ntop = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = "\n".join(subprocess.run(["nvdisasm", "-g", "-c", c], capture_output=True, text=True).stdout
                for c in sorted(glob.glob(tmp + "/*.cubin")))      # one cubin per translation unit
# instruction order -> (line, inlined-from chain ignored)
lines = []; cur = None; inside = False
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l: inside = ksub in l; continue
    if not inside: continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
blocks = []; b = None
for r in rows:
    if r and r[0] == "Kernel Name": b = {"name": r[1], "rows": []}; blocks.append(b)
    elif b is not None: b["rows"].append(r)
hdr = blocks[0]["rows"][0]; data = blocks[0]["rows"][1:]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
print("sass rows", len(data), "disasm instrs", len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for r, ln in zip(data, lines):
    agg[ln][0] += int(r[iex] or 0); agg[ln][1] += int(r[isamp] or 0)
srcfile = sys.argv[5] if len(sys.argv) > 5 else "jaxovercooked_b200/csrc/oc_kernels.cuh"
srcs = {}
def text(key):
    if not key: return ""
    f, n = key
    if f not in srcs:
        cand = [os.path.join(os.path.dirname(srcfile), f), srcfile]
        srcs[f] = next((open(c).read().splitlines() for c in cand if os.path.basename(c) == f and os.path.exists(c)), [])
    return srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp-instr", tot, "samples", ts)
for ln, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print(f"{ex:>9} {100*ex/tot:5.1f}%  samp {sm:>4} {100*sm/max(ts,1):5.1f}%  {ln[0] if ln else '?'}:{ln[1] if ln else 0}: {text(ln)}")

# ---- region totals (line ranges of the main source file given as extra args "name:lo-hi")
main = os.path.basename(srcfile)
regs = [a for a in sys.argv[6:] if ":" in a]
for rg in regs:
    name, lh = rg.split(":"); lo, hi = map(int, lh.split("-"))
    sel = [v for ln, v in agg.items() if ln and ln[0] == main and lo <= ln[1] <= hi]
    ex = sum(v[0] for v in sel); sm = sum(v[1] for v in sel)
    print(f"REGION {name:14s} L{lo}-{hi}: instr {ex:>10} {100*ex/tot:5.1f}%  samples {sm:>5} {100*sm/max(ts,1):5.1f}%")
