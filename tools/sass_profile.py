"""Instruction mix and hot regions from an `ncu --page source --csv` export (SASS view):
   python tools/sass_profile.py gpurun_out/src_k_x.csv [region_size]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def f(r, n):
    try: return float(r[col[n]])
    except ValueError: return 0.0
tot_i = sum(f(r, "Instructions Executed") for r in data); tot_s = sum(f(r, "# Samples") for r in data)
print("SASS lines %d, warp-instructions %.4g, samples %d" % (len(data), tot_i, tot_s))
ops = collections.Counter(); ops_s = collections.Counter()
for r in data:
    toks = r[col["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += f(r, "Instructions Executed"); ops_s[op] += f(r, "# Samples")
print("top opcodes (share of executed warp-instructions | share of stall samples):")
for op, c in ops.most_common(28):
    print("  %-10s %5.1f%% | %5.1f%%" % (op, 100 * c / tot_i, 100 * ops_s[op] / max(tot_s, 1)))
reg = int(sys.argv[2]) if len(sys.argv) > 2 else 64
print("regions of %d SASS lines (executed share | sample share | first instruction):" % reg)
for i in range(0, len(data), reg):
    blk = data[i:i + reg]
    ci = sum(f(r, "Instructions Executed") for r in blk); cs = sum(f(r, "# Samples") for r in blk)
    print("  %5d %5.1f%% | %5.1f%% | %s" % (i, 100 * ci / tot_i, 100 * cs / max(tot_s, 1), blk[0][col["Source"]].strip()[:60]))
