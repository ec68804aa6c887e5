"""One row per kernel from an `ncu --page raw --csv` export: python tools/ncu_summary.py raw.csv [summary.json]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] == "": return default
    try: return float(r[i].replace(",", ""))
    except ValueError: return default
def unit(name): return units[col[name]] if name in col else ""
def to(v, u, want):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}
    return v * f.get(u, 1) / f.get(want, 1)
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not stalls:
    stalls = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
out = []
tot = sum(to(g(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"), "us") for r in data)
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    t = to(g(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"), "us")
    rd = to(g(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"), "Mbyte")
    wr = to(g(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"), "Mbyte")
    st = sorted(((g(r, h, 0.0), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "").replace(".ratio", "")) for h in stalls), reverse=True)
    ssum = sum(v for v, _ in st) or 1.0
    out.append(dict(kernel=name, us=t, share=100 * t / tot, grid=r[col["Grid Size"]], block=r[col["Block Size"]],
                    regs=g(r, "launch__registers_per_thread"), dram_rd=rd, dram_wr=wr,
                    dram_pct=g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                    sm_pct=g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                    lsu_pct=g(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                    l2_pct=g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                    warps=g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    inst=g(r, "smsp__inst_executed.sum"), issue=g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    adu_pct=g(r, "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active"),
                    fma_pct=g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                    alu_pct=g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                    stalls=", ".join("%s %.0f%%" % (n, 100 * v / ssum) for v, n in st[:4])))
print("| kernel | µs (share) | grid×block | regs | DRAM rd/wr MB | DRAM % | L2 % | LSU % | SM % | warps % | warp-inst | issue % | pipes fma/alu/adu % | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for o in out:
    print("| %(kernel)s | %(us).1f (%(share).0f%%) | %(grid)s×%(block)s | %(regs).0f | %(dram_rd).0f / %(dram_wr).0f | %(dram_pct).0f | %(l2_pct).0f | %(lsu_pct).0f | %(sm_pct).0f | %(warps).0f | %(inst).3g | %(issue).0f | %(fma_pct).0f / %(alu_pct).0f / %(adu_pct).0f | %(stalls)s |" % o)
print("\nTotal: %.1f µs" % tot)
if len(sys.argv) > 2:
    import json
    json.dump({"kernels": out, "total_us": tot}, open(sys.argv[2], "w"), indent=1)
if len(sys.argv) > 3:
    # traffic of the roofline kernel (bench.py reads it): DRAM bytes of ONE k_preprocess launch of this capture
    import json
    pre = next((o for o in out if o["kernel"].startswith("k_preprocess")), None)
    if pre:
        json.dump({"source": "%s (ncu --set full, one launch of %s, written by tools/ncu_summary.py)" % (sys.argv[1], pre["kernel"]),
                   "preprocess_dram_bytes_per_launch": (pre["dram_rd"] + pre["dram_wr"]) * 1e6,
                   "preprocess_dram_read_MB": pre["dram_rd"], "preprocess_dram_write_MB": pre["dram_wr"]},
                  open(sys.argv[3], "w"), indent=1)
