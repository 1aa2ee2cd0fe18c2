"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).

  python tools/summarize_ncu.py r01     # reads gpurun_out/launches_bench.csv and gpurun_out/prof_*.ncu-rep
"""
import collections, csv, glob, io, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
lines = [f"# ncu summary {tag} (C3: 1 997 120 faces / 200 views 1080p, 1x B200)", ""]

# ---- launch list: share of the step per kernel (cold-cache, serialised: compare SHARES, not absolutes) ----
lp = os.path.join(ROOT, "gpurun_out", f"{tag}_launches_bench.csv")
if not os.path.exists(lp):
    lp = os.path.join(ROOT, "gpurun_out", "launches_bench.csv")
if os.path.exists(lp):
    txt = open(lp, errors="replace").read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        name = name.replace("b2::<unnamed>::", "").replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    lines += ["## Launch list of `python bench.py --steps 2 --warmup 1 --no-verify --no-e2e --no-cpu-baseline` (3 passes of the hot path)",
              "", f"total kernel time {total:.1f} ms over {sum(a[0] for a in agg.values())} launches", "",
              "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        lines.append(f"| `{n[:70]}` | {c} | {ms:.3f} | {100 * ms / total:.1f}% |")
    lines.append("")
    import shutil
    if os.path.abspath(lp) != os.path.abspath(os.path.join(out_dir, f"{tag}_launches_bench.csv")):
        shutil.copy(lp, os.path.join(out_dir, f"{tag}_launches_bench.csv"))

# ---- full captures ----
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
traffic = {}
lines += ["## `ncu --set full --clock-control none` captures (one launch each, `tools/run_pipeline.py C3 1`)", ""]
reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_prof_*.ncu-rep"))) or sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep")))
for rep in reps:
    k = os.path.basename(rep)[:-8].split("prof_")[1]
    try:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=120).stdout
    except Exception as e:
        lines.append(f"### {k}: could not read ({e})")
        continue
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else k
    lines += [f"### {k}  (`{kn[:90]}`)", "", "| metric | value | unit |", "|---|---:|---|"]
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            d[w] = (vals[i], units[i])
            lines.append(f"| {w} | {vals[i]} | {units[i]} |")
    def tobytes(key):
        if key not in d:
            return 0.0
        v, u = d[key]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    tb = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    traffic[k] = tb
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
            try:
                stalls.append((float(vals[i].replace(",", "")), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(stalls, reverse=True)[:5])
    lines += ["", f"dram traffic per launch: {tb / 1e6:.1f} MB; top stall reasons (pc sampling): {top}", ""]
open(os.path.join(out_dir, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(out_dir, f"{tag}_ncu_traffic.json"), "w"), indent=1)
print("\n".join(lines[:60]))
