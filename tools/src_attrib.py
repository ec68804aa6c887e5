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
ins = []   # (file, line) of every instruction; helpers inlined from other files keep their own file:line
cur = None
base = srcfile.split("/")[-1]
srcdir = "/".join(srcfile.split("/")[:-1])
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
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
import os
_src = {}
def text(key):
    if not key:
        return "?"
    f, ln = key
    if f not in _src:
        path = os.path.join(srcdir, f)
        _src[f] = open(path).read().split("\n") if os.path.exists(path) else None
    body = _src[f][ln - 1].strip()[:96] if _src[f] and ln <= len(_src[f]) else "(library header)"
    return ("" if f == base else f + ": ") + body
print("SASS %d (disasm %d), warp-instructions %.4g, samples %d" % (len(data), len(ins), ti, ts))
for key, c in sorted(by_i.items(), key=lambda x: -x[1])[:top]:
    print("%5s %5.1f%% instr | %5.1f%% samples | %s" % (key[1] if key else "?", 100 * c / ti, 100 * by_s[key] / max(ts, 1), text(key)))
# per file, and for this file per function-sized block of 25 lines, so that spread-out costs add up
byfile = collections.Counter()
for key, c in by_i.items():
    byfile[key[0] if key else "?"] += c
print("by file: " + ", ".join("%s %.1f%%" % (f, 100 * c / ti) for f, c in byfile.most_common(6)))
