"""Summarise gpurun_out/<tag>/ (ncu launch lists + --set full captures + bench lines) into profiles/<tag>_*.{csv,md,json}.

    python tools/ncu_summary.py r1a
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]


def raw_rows(rep):
    # either a report (read through ncu) or the `--page raw --csv` export tools/gpu_round.sh keeps (reports exceed gpurun's 64 MiB)
    out = open(rep, errors="replace").read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [{h: (r[i], units[i]) for i, h in enumerate(hdr)} for r in rows[2:]]


def main(tag):
    src, dst = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    md = [f"# ncu summary `{tag}` (B200, `--clock-control none`)\n"]
    for f in sorted(os.listdir(src)):
        p = os.path.join(src, f)
        if f.startswith("launches_") and f.endswith(".csv"):
            rows = [r for r in csv.reader(open(p, errors="replace")) if len(r) > 14 and r[0].isdigit()]
            agg = {}
            for r in rows:
                k = r[4].split("(")[0][:90]
                a = agg.setdefault(k, [0, 0.0])
                a[0] += 1; a[1] += float(r[14].replace(",", "")) * (1e-3 if r[13] == "ns" else 1.0)
            tot = sum(a[1] for a in agg.values())
            with open(os.path.join(dst, f"{tag}_{f}"), "w") as o:
                o.write("kernel,launches,total_us,mean_us,share\n")
                for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    o.write(f"\"{k}\",{a[0]},{a[1]:.1f},{a[1] / a[0]:.2f},{a[1] / tot:.4f}\n")
            md.append(f"## launch list `{f}` ({len(rows)} launches, gpu__time_duration)\n")
            md.append("| kernel | launches | mean µs | share |\n|---|---|---|---|")
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                md.append(f"| `{k}` | {a[0]} | {a[1] / a[0]:.2f} | {100 * a[1] / tot:.1f} % |")
            md.append("")
        if f.endswith(".ncu-rep") or f.endswith(".raw.csv"):
            recs = raw_rows(p)
            md.append(f"## `--set full` capture `{f}`\n")
            traffic = {}
            for i, r in enumerate(recs):
                name = r["Kernel Name"][0].split("(")[0]
                md.append(f"### launch {i}: `{name}`\n\n| metric | value | unit |\n|---|---|---|")
                for k in KEYS:
                    if k in r:
                        md.append(f"| {k} | {r[k][0]} | {r[k][1]} |")
                md.append("")
                def b(k):
                    v, u = r[k]
                    return float(v) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[u]
                traffic.setdefault(name.replace("void ", "").split("<")[0], []).append(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
            wl = f.replace("prof_", "").replace(".ncu-rep", "").replace(".raw.csv", "")
            json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(os.path.join(dst, f"traffic_{wl}.json"), "w"))
        if f.endswith(".funcs.txt") or f.endswith(".lines.txt") or f.startswith("trace_"):
            txt = open(p, errors="replace").read()
            open(os.path.join(dst, f"{tag}_{f}"), "w").write(txt)
            if f.endswith(".funcs.txt"):
                md.append(f"## per-function attribution `{f}` (tools/ncu_funcs.py)\n\n```\n" + "\n".join(txt.splitlines()[:24]) + "\n```\n")
            if f.startswith("trace_"):
                md.append(f"## per-environment phase timeline `{f}` (tools/step_trace.py)\n\n```\n" + txt + "\n```\n")
        if f.startswith("bench_") and f.endswith(".json"):
            lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
            if lines:
                open(os.path.join(dst, f"{tag}_{f}"), "w").write(lines[-1] + "\n")
    open(os.path.join(dst, f"{tag}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md)[:6000])


if __name__ == "__main__":
    main(sys.argv[1])
