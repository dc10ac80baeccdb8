"""Turns the raw evidence collected by scripts/collect_profiles.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); Pdir = os.path.join(ROOT, "profiles")
os.makedirs(Pdir, exist_ok=True)
R = sys.argv[1] if len(sys.argv) > 1 else "r1"

def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))

def ncu_details(rep):
    return subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout

# ---- launch list ----
lines = open(os.path.join(G, f"launches_{R}.csv")).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(r["Kernel Name"].split("(")[0].replace("void ", ""), []).append(float(r["Metric Value"]))
ours = {k: v for k, v in agg.items() if k.startswith("strict::") and "fp64_peak" not in k}
tot = sum(sum(v) for v in ours.values())
shutil.copy(os.path.join(G, f"launches_{R}.csv"), os.path.join(Pdir, f"{R}_launches_bench.csv"))
bench = json.loads(open(os.path.join(G, f"bench_{R}.json")).read().strip().split("\n")[-1])
km = bench["kernel_ms"]; kms = sum(km.values())
with open(os.path.join(Pdir, f"{R}_launches_bench_summary.md"), "w") as f:
    f.write(f"# {R}: ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-lbfgs`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_%s.csv python bench.py ...`\n" % R)
    f.write(f"(raw: `profiles/{R}_launches_bench.csv`).  Per-launch times under ncu are cold-cache and serialised, so the SHARES are compared\n")
    f.write("with the CUDA-event breakdown that the un-profiled `bench.py` run prints (`kernel_ms`, right column).\n\n")
    f.write("| kernel | launches | mean µs (ncu) | share (ncu) | share (CUDA events, bench.py) |\n|---|---|---|---|---|\n")
    ev = {"strict::k_pose_table": km["k_pose_table"], "strict::k_outer<0, 0>": km["k_outer"], "strict::k_finalize": km["k_finalize"]}
    for k, v in agg.items():
        if k in ours:
            e = ev.get(k)
            es = f"{100*e/kms:.1f} %" if e is not None else ("%.1f %% (k_compact + k_gsip)" % (100 * km["k_compact+k_gsip"] / kms) if "gsip" in k else "(with k_gsip)")
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/tot:.1f} % | {es} |\n")
        else:
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | not part of the step (L2 flush / FP64 peak micro-benchmark) | |\n")

# ---- full captures ----
def summarise(rep, name, kernel_regex):
    raw, units = ncu_raw(rep)
    det = ncu_details(rep)
    keep = ["Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Executed Ipc Active", "Issue Slots Busy",
            "No Eligible", "Active Warps Per Scheduler", "Eligible Warps Per Scheduler", "Avg. Active Threads Per Warp", "Block Limit Registers",
            "Dynamic Shared Memory Per Block", "Waves Per SM", "Grid Size", "Block Size"]
    lines_ = [l for l in det.splitlines() if any(k in l for k in keep)]
    g = lambda k: float(raw[k]) if k in raw and raw[k] not in ("", "n/a") else None
    dr, dw = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    ur, uw = units.get("dram__bytes_read.sum", ""), units.get("dram__bytes_write.sum", "")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = (dr or 0) * scale.get(ur, 1) + (dw or 0) * scale.get(uw, 1)
    stalls = sorted(((float(v), k) for k, v in raw.items() if "smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)[:8]
    pipes = {k.split("sm__inst_executed_pipe_")[1].split(".")[0]: float(v) for k, v in raw.items()
             if k.startswith("sm__inst_executed_pipe_") and k.endswith(".avg.pct_of_peak_sustained_active") and v not in ("", "n/a") and float(v) > 0.5}
    with open(os.path.join(Pdir, f"{R}_{name}_ncu_summary.md"), "w") as f:
        f.write(f"# {R}: `ncu --set full --clock-control none --import-source on -k regex:{kernel_regex} -s 2 -c 1` of `python scripts/prof_step.py`\n\n")
        f.write("Workload: config 2 (star, 8-piece MINCO, 200 000 query points), strict build.  Numbers under the profiler are not bench values.\n\n```\n")
        f.write("\n".join(l.rstrip() for l in lines_) + "\n```\n\n")
        f.write(f"DRAM traffic of this launch: read {dr} {ur}, write {dw} {uw}  => {traffic/1e6:.2f} MB "
                "(algorithmic: 3.2 MB of query points + the 7 KB trajectory blob per CTA from L2)\n\n")
        f.write("Pipe utilisation (`sm__inst_executed_pipe_*.avg.pct_of_peak_sustained_active`, > 0.5 %): "
                + ", ".join(f"{k} {v:.1f} %" for k, v in sorted(pipes.items(), key=lambda kv: -kv[1])) + "\n\n")
        f.write("Executed warp instructions: %s\n\n" % raw.get("smsp__inst_executed.sum"))
        f.write("Top warp stall reasons (`smsp__average_warps_issue_stalled_*_per_issue_active.ratio`):\n\n")
        for v, k in stalls:
            f.write(f"* {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio','')}: {v:.2f}\n")
        f.write("\nReading: DRAM and tensor pipes are idle; the FP64 pipe and the issue slots are the busy resources (FP64-compute / latency bound,\n"
                "as DESIGN.md §3 predicts).\n")
    return traffic

t_outer = summarise(os.path.join(G, f"prof_outer_{R}.ncu-rep"), "k_outer", "k_outer")
summarise(os.path.join(G, f"prof_gsip_{R}.ncu-rep"), "k_gsip", "k_gsip")
json.dump({"kernel": "k_outer", "dram_bytes_per_launch": t_outer, "source": f"profiles/{R}_k_outer_ncu_summary.md (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture)"},
          open(os.path.join(Pdir, f"{R}_k_outer_traffic.json"), "w"), indent=1)
# ---- bench / configs / batch records ----
for src, dst in ((f"bench_{R}.json", f"{R}_bench.json"), (f"bench_reference_{R}.json", f"{R}_bench_reference.json"), (f"configs_{R}.jsonl", f"{R}_configs_1_to_4.jsonl"),
                 (f"batch_1gpu_{R}.json", f"{R}_batch_1gpu.json"), (f"pytest_gpu_{R}.log", f"{R}_pytest_gpu.log")):
    if os.path.exists(os.path.join(G, src)): shutil.copy(os.path.join(G, src), os.path.join(Pdir, dst))
print("profiles written:", sorted(os.listdir(Pdir)))
