"""Regenerate the ncu evidence bench.py's roofline block reads (run ON the GPU box, one GPU):

    python tools/ncu_profile.py --git-hash $(git rev-parse --short HEAD) [--workloads breaktime cornell] [--tag r2]

For every workload it captures, with `ncu --set full --clock-control none --import-source on`, the bounce-1 launches of
the extend kernel (nearest-hit trace), the shadow-connect kernel (any-hit trace) and the shade kernel of one short
render (tools/prof_run.py, CUDA graphs off so every launch is visible), reads the reports back with `ncu -i ... --page
raw --csv`, and writes

    profiles/kernel_profiles.json        per workload and kernel: duration, warp instructions (per ray / per hit),
                                         lanes per instruction, issue-slot utilisation, pipe utilisation, L1 / L2 hit
                                         rates, DRAM bytes (per ray / per hit) — with the git hash of the capture
    profiles/<tag>_<kernel>_<workload>_ncu_summary.txt   the tracked metrics + the heaviest SASS basic blocks

The .ncu-rep files land in gpurun_out/ (scratch).  Numbers printed by a run under ncu are never bench values.
"""
import argparse
import csv
import datetime
import io
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
PROFILES = os.path.join(REPO, "profiles")

KERNELS = {
    # name: (ncu -k regex, launches to skip, what the unit of work is)
    "extend": ("wf_trace", 2, "rays"),   # launches in order: extend b0, shadow b0, extend b1, shadow b1, ...
    "shadow": ("wf_trace", 3, "rays"),
    "shade": ("wf_shade", 1, "hits"),
}


def run_capture(workload, spp, kernel, tag):
    regex, skip, _ = KERNELS[kernel]
    rep = os.path.join(OUT, f"{tag}_{kernel}_{workload}")
    env = dict(os.environ, RPT_GRAPHS="0", RPT_LOG_QUEUES="1")
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{regex}", "-s", str(skip), "-c", "1", "-f", "-o", rep,
           sys.executable, os.path.join(REPO, "tools", "prof_run.py"), workload, str(spp)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if p.returncode != 0:
        raise SystemExit(f"ncu failed for {kernel}/{workload}:\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    # RPT_LOG_QUEUES: "[rpt] bounce 1: extend traced N rays -> H hits, M misses; S shadow rays; P paths go on"
    m = re.search(r"bounce 1: extend traced (\d+) rays -> (\d+) hits, (\d+) misses; (\d+) shadow rays", p.stderr)
    if not m:
        raise SystemExit("queue log of bounce 1 not found in the render's stderr:\n" + p.stderr[-2000:])
    rays, hits, _misses, shadow = (int(g) for g in m.groups())
    return rep + ".ncu-rep", {"extend": rays, "shadow": shadow, "shade": hits}[kernel]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, row = rows[0], rows[1], rows[2]

    def metric(name):
        c = head.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[c], 1.0)
        return float(row[c].replace(",", "")) * scale

    return metric, row[head.index("Kernel Name")]


def summarize(rep, kernel, workload, units_of_work, unit_name, tag, git_hash):
    metric, kernel_name = raw_metrics(rep)
    dram = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
    inst = metric("smsp__inst_executed.sum")
    entry = {
        "kernel_name": kernel_name, "launch": f"bounce 1 of one {workload} wave", unit_name: units_of_work,
        "duration_ms": metric("gpu__time_duration.sum"),
        "warp_instructions": inst, f"warp_instructions_per_{unit_name[:-1]}": inst / units_of_work,
        "lanes_per_instruction": metric("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct_of_peak": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": metric("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": metric("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "l1_hit_pct": metric("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": metric("lts__t_sector_hit_rate.pct"),
        "dram_bytes": dram, f"dram_bytes_per_{unit_name[:-1]}": dram / units_of_work,
        "registers_per_thread": metric("launch__registers_per_thread"),
        "warps_active_pct": metric("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "long_scoreboard_per_issue": metric("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "source": os.path.basename(rep) + " (ncu --set full --clock-control none)",
    }
    # human-readable summary next to it
    csv_path = rep.replace(".ncu-rep", ".csv")
    with open(csv_path, "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    text = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_summary.py"), csv_path], capture_output=True, text=True).stdout
    blocks = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_blocks.py"), rep, {"shade": "wf_shade"}.get(kernel, "wf_trace"), "14"],
                            capture_output=True, text=True).stdout
    with open(os.path.join(PROFILES, f"{tag}_{kernel}_{workload}_ncu_summary.txt"), "w") as f:
        f.write(f"# {tag} {kernel_name[:100]}\n# {entry['launch']}: {units_of_work} {unit_name}; git {git_hash}; ncu --set full --clock-control none\n")
        f.write("\n".join(line[:200] for line in text.splitlines() if ".max." not in line and ".min." not in line and ".sum.p" not in line) + "\n")
        f.write("\n# heaviest SASS basic blocks (share of warp instructions | instructions x executions | lanes enabled | stall samples | opcode mix)\n")
        f.write("\n".join(line[:220] for line in blocks.splitlines()) + "\n")
    return entry


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--git-hash", required=True)
    ap.add_argument("--workloads", nargs="+", default=["breaktime", "cornell"])
    ap.add_argument("--kernels", nargs="+", default=list(KERNELS))
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--tag", default="r2")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(PROFILES, "kernel_profiles.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table["git"] = args.git_hash
    table["captured"] = datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ")
    table["how"] = "python tools/ncu_profile.py (ncu --set full --clock-control none, bounce-1 launches of tools/prof_run.py, graphs off)"
    for workload in args.workloads:
        table.setdefault(workload, {})
        for kernel in args.kernels:
            rep, units = run_capture(workload, args.spp, kernel, args.tag)
            table[workload][kernel] = summarize(rep, kernel, workload, units, KERNELS[kernel][2], args.tag, args.git_hash)
            print(workload, kernel, json.dumps(table[workload][kernel])[:400], flush=True)
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
