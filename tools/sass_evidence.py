"""SASS evidence of the built kernels (offline; needs cuobjdump + nvdisasm, no GPU):
   python tools/sass_evidence.py [outdir=profiles]
For every kernel of wgpu-3dgs-viewer-app_b200/_build/*.o: instruction count, opcode histogram, and every line carrying one of
the mnemonics that prove the Blackwell-native paths (TMA bulk copy + mbarrier, cluster barrier / DSMEM, votes and matches,
MUFU, global reductions), each with the source line it comes from (-lineinfo)."""
import collections, glob, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, sys.argv[1] if len(sys.argv) > 1 else "profiles")
KEY = re.compile(r"\b(UBLKCP|SYNCS|UCGABAR|MAPA|MATCH|VOTE|MUFU|REDG|RED|ATOMS|ATOMG|ELECT|FFMA2|LDGSTS|UTMA\w*|BAR|CCTL|ERRBAR|MEMBAR|FENCE|ACQBULK|ST\.E\.\w*\.?CLUSTER|LD\.E)\b")
WANT = {"k_preprocess": "ILi2ELi1ELb1", "k_sort_pass": "k_sort_passILi5EE", "k_sort_pass_wide": "ILi11ELi8ELb1", "k_bin": "", "k_tile_finish": "", "k_composite": "ILb0ELb0",
        "k_eval_mask": "", "k_postprocess": "", "k_paint_query_texture": "", "k_query_hits": "", "k_sort_hist": ""}


def main():
    os.makedirs(OUT, exist_ok=True)
    for obj in sorted(glob.glob(os.path.join(ROOT, "wgpu-3dgs-viewer-app_b200", "_build", "*.o"))):
        with tempfile.TemporaryDirectory() as td:
            subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, capture_output=True)
            for cubin in glob.glob(os.path.join(td, "*.cubin")):
                txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
                parts = re.split(r"\n\s*\.section\s+\.text\.", txt)
                for part in parts[1:]:
                    name = part.split(",", 1)[0]
                    short = next((k for k in WANT if k.replace("_wide", "") in name and WANT[k] in name), None)
                    if not short:
                        continue
                    cur, ops, keys, n = "?", collections.Counter(), [], 0
                    for line in part.split("\n"):
                        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
                        if m:
                            cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
                            continue
                        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
                        if not m:
                            continue
                        n += 1
                        ins = m.group(2).strip()
                        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
                        ops[op] += 1
                        if KEY.search(ins):
                            keys.append("  /*%s*/ %-70s // %s" % (m.group(1), ins[:70], cur))
                    path = os.path.join(OUT, "r2_sass_%s.txt" % short)
                    with open(path, "w") as f:
                        f.write("# %s\n# from %s (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo), %d SASS instructions\n" % (name, os.path.basename(obj), n))
                        f.write("# opcode histogram: " + ", ".join("%s %d" % kv for kv in ops.most_common(24)) + "\n")
                        f.write("# lines with TMA / mbarrier / cluster / vote / match / MUFU / reduction mnemonics (source line from -lineinfo):\n")
                        seen = collections.Counter()
                        for k in keys:
                            mn = KEY.search(k).group(1)
                            seen[mn] += 1
                            if seen[mn] <= 12:
                                f.write(k + "\n")
                        f.write("# totals: " + ", ".join("%s x%d" % kv for kv in seen.most_common()) + "\n")
                    print("wrote", os.path.relpath(path, ROOT), n, dict(seen.most_common(8)))


if __name__ == "__main__":
    main()
