"""Attributes an `ncu --page source --csv` SASS export to SOURCE LINES through `nvdisasm -g` of the same cubin:
   python tools/src_attrib.py <src.csv> <file.sass from `nvdisasm -g -c`> <mangled-name substring> <source file> [top]
(instructions are matched by position: both listings are in address order)"""
import csv, re, sys, collections
csvp, sassp, fn, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(csvp)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
lines = open(sassp).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and fn in l)
ins = []   # (line number in the kernel's own file, opcode text)
cur = None
base = srcfile.split("/")[-1]
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, n, rest = m.group(1), int(m.group(2)), m.group(3)
        if not f.endswith(base):
            # inlined library code: attribute to the innermost frame of our file
            mm = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
            own = [int(b) for a, b in mm if a.endswith(base)]
            cur = own[0] if own else cur
        else:
            cur = n
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        ins.append(cur)
assert len(ins) >= len(data), (len(ins), len(data))
def f(r, n):
    try: return float(r[col[n]])
    except ValueError: return 0.0
by_i = collections.Counter(); by_s = collections.Counter()
for k, r in enumerate(data):
    by_i[ins[k]] += f(r, "Instructions Executed"); by_s[ins[k]] += f(r, "# Samples")
ti = sum(by_i.values()); ts = sum(by_s.values())
src = open(srcfile).read().split("\n")
print("SASS %d (disasm %d), warp-instructions %.4g, samples %d" % (len(data), len(ins), ti, ts))
for ln, c in sorted(by_i.items(), key=lambda x: -x[1])[:top]:
    print("%5s %5.1f%% instr | %5.1f%% samples | %s" % (ln, 100 * c / ti, 100 * by_s[ln] / max(ts, 1), (src[ln - 1].strip()[:100] if ln else "?")))
