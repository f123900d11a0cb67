#!/usr/bin/env python
"""Attribute ncu warp-stall samples of one kernel to CUDA source lines (no GPU needed):
joins `ncu --page source --csv` (per-SASS-instruction samples) with `nvdisasm --print-line-info`
of the in-tree library by instruction order.  Usage: tools/ncu_lines.py <rep> <kernel-substring>"""
import collections, csv, io, re, subprocess, sys, os, tempfile, glob
rep, key = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "opengjk_b200", "libopengjk_b200.so")], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "--print-line-info"] + glob.glob(os.path.join(tmp, "*.cubin")), capture_output=True, text=True).stdout
# instructions of the kernel with their (file, line, inlined-at chain top)
ins, cur, line = [], None, None
for l in sass.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur = m.group(1); line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and key in cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        op = re.sub(r"/\*.*?\*/", "", l).strip().split()[0:2]
        ins.append((line, " ".join(op)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
prof = []
for r in rows[2:]:
    try:
        prof.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])))
    except (ValueError, IndexError):
        pass
print(f"sass instructions: binary {len(ins)}  profile {len(prof)}")
n = min(len(ins), len(prof))
mism = sum(1 for i in range(n) if ins[i][1].split()[0].split('.')[0].lstrip('@!P0123456789 ') not in prof[i][0])
print("opcode mismatches:", mism)
by = collections.defaultdict(lambda: [0, 0, 0])
for i in range(n):
    by[ins[i][0]][0] += prof[i][1]; by[ins[i][0]][1] += prof[i][2]; by[ins[i][0]][2] += prof[i][3]
tot = sum(v[0] for v in by.values()); toti = sum(v[1] for v in by.values())
srcs = {}
for (k, v) in sorted(by.items(), key=lambda kv: -kv[1][0])[:40]:
    if k is None: continue
    f, ln = k
    if f not in srcs:
        pth = os.path.join(ROOT, "opengjk_b200", "csrc", f)
        srcs[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
    text = srcs[f][ln - 1].strip()[:80] if ln - 1 < len(srcs[f]) else ""
    print(f"{100*v[0]/tot:5.1f}% samples {100*v[1]/toti:5.1f}% instr lanes {v[2]/max(v[1],1):4.1f}  {f}:{ln}  {text}")
