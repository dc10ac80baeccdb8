"""Turn the `ncu --page raw --csv` export of one k_outer launch into the record bench.py reports its roofline from.

    python scripts/roofline_from_ncu.py gpurun_out/prof_<tag>_raw.csv profiles/r2_k_outer_roofline.json [--note "..."]

flop_per_launch = (DADD + DMUL + 2 DFMA) thread instructions of the launch, from
smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum.per_cycle_elapsed x smsp cycles elapsed — the executed FP64
arithmetic, FMA counted as two (DSETP, conversions and the MUFU seeds of sqrt / division are not counted).
dram_bytes_per_launch = dram__bytes_read.sum + dram__bytes_write.sum.  The capture is of the config-2 workload
(scripts/prof_step.py: star, 8 pieces, 200 000 points, strict build); numbers taken under the profiler are not bench values,
bench.py divides flop_per_launch by the kernel time it measures itself with CUDA events.
"""
import csv
import json
import sys


def to_bytes(val, unit):
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    return float(val) * mult


def main():
    raw, out = sys.argv[1], sys.argv[2]
    note = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--note" else ""
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    f = lambda k: float(m[k][0])
    cycles = f("smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in m else f("sm__cycles_elapsed.avg")
    dadd = f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cycles
    dmul = f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cycles
    dfma = f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cycles
    rec = {
        "kernel": m["Kernel Name"][0] if "Kernel Name" in m else "k_outer",
        "workload": "config 2: star, 8-piece MINCO, 200000 query points, strict build (scripts/prof_step.py)",
        "flop_per_launch": dadd + dmul + 2.0 * dfma,
        "thread_inst": {"dadd": dadd, "dmul": dmul, "dfma": dfma},
        "cycles_elapsed": cycles,
        "duration_ms_under_ncu": to_bytes(1, "byte") * f("gpu__time_duration.sum") * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}[m["gpu__time_duration.sum"][1]],
        "dram_bytes_per_launch": to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"]),
        "fp64_pipe_active_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_elapsed_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": f("smsp__inst_executed.sum"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "source": raw,
        "note": note,
    }
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
