"""Aggregate an ncu `--page source --csv` SASS dump by CUDA source line using nvdisasm line info.
usage: ncu_by_line.py <ncu_source.csv> <nvdisasm.sass> <mangled-kernel-prefix> [source-file-to-annotate]"""
import collections, csv, re, sys
ncu_csv, sass, kern = sys.argv[1:4]
src_path = sys.argv[4] if len(sys.argv) > 4 else None
# nvdisasm: offset -> (file, line)
off2line = {}
cur = None; inside = False
for ln in open(sass, errors="ignore"):
    if ln.startswith(".text." + kern):
        inside = True; continue
    if inside and ln.startswith("//---------------------") : break
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
    if m: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); ist = hdr.index("# Samples")
base = None
byline = collections.Counter(); samp = collections.Counter(); tot = 0; fp64 = collections.Counter()
for r in rows[2:]:
    try: addr = int(r[0], 16); n = int(r[ia]); s = int(r[ist])
    except Exception: continue
    if base is None: base = addr
    key = off2line.get(addr - base, ("?", 0))
    byline[key] += n; samp[key] += s; tot += n
    if re.search(r'\b(DFMA|DMUL|DADD|DSETP)\b', r[1]): fp64[key] += n
src = open(src_path).read().split("\n") if src_path else None
print(f"total warp-instructions {tot}")
for (f, l), n in byline.most_common(45):
    text = src[l - 1].strip()[:90] if src and f.endswith(src_path.split("/")[-1]) and 0 < l <= len(src) else ""
    print(f"{f}:{l:<5d} {n:12d} {100*n/tot:6.2f}%  fp64 {100*fp64[(f,l)]/max(n,1):5.1f}%  samples {samp[(f,l)]:6d} | {text}")

# per-opcode per-line breakdown for selected opcodes
import os
ops = os.environ.get("NCU_OPS")
if ops:
    ops = ops.split(",")
    per = {o: collections.Counter() for o in ops}
    base = None
    for r in rows[2:]:
        try: addr = int(r[0], 16); n = int(r[ia])
        except Exception: continue
        if base is None: base = addr
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[1])
        op = m.group(2) if m else ""
        if op in per: per[op][off2line.get(addr - base, ("?", 0))] += n
    for o in ops:
        print(f"--- {o}: total {sum(per[o].values())}")
        for (f, l), n in per[o].most_common(12):
            text = src[l - 1].strip()[:80] if src and f.endswith(src_path.split('/')[-1]) and 0 < l <= len(src) else ""
            print(f"   {f}:{l:<5d} {n:12d} | {text}")
