"""Turn ncu reports / launch lists brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_ncu.py <tag> <mode> <gpurun_out dir>
"""
import collections, csv, json, os, subprocess, sys

tag, mode, src = sys.argv[1], sys.argv[2], sys.argv[3]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__icc_request_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

rep = os.path.join(src, f"render_{mode}.ncu-rep")
summary = {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    summary = {"kernel": d["Kernel Name"][0]}
    for k in KEYS:
        if k in d:
            summary[k] = {"value": d[k][0], "unit": d[k][1]}
    def mb(k):
        v, u = d[k]
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    summary["dram_bytes_per_launch"] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    with open(os.path.join(out, f"{tag}_render_{mode}_ncu_full.json"), "w") as f:
        json.dump(summary, f, indent=1)
    # bench.py reads this one for roofline.traffic
    pj = os.path.join(out, "render_kernel_ncu.json")
    allm = json.load(open(pj)) if os.path.exists(pj) else {}
    allm[mode] = {"kernel": summary["kernel"], "dram_bytes_per_launch": summary["dram_bytes_per_launch"],
                  "source": f"profiles/{tag}_render_{mode}_ncu_full.json"}
    json.dump(allm, open(pj, "w"), indent=1)
    print("full capture:", summary["kernel"], summary.get("gpu__time_duration.sum"))

ll = os.path.join(src, f"launches_{mode}.csv")
if os.path.exists(ll):
    lines = [l for l in open(ll) if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        k = row["Kernel Name"]
        order.append((k, v))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, f"{tag}_launches_{mode}.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : bench.py --steps 2 --warmup 3 --mode {mode}\n")
        f.write(f"# per-launch times are cold-cache and serialised: compare SHARES.  total {tot:.3f} ms over {len(order)} launches\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:10.3f} ms {100 * t / tot:6.2f}%  n={n:4d}  {k[:120]}\n")
    print("launch list:", len(order), "launches,", f"{tot:.2f} ms")
