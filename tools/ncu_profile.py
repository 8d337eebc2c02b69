"""Regenerate the ncu evidence bench.py's roofline block reads (run ON the GPU box, one GPU):

    python tools/ncu_profile.py --git-hash $(git rev-parse --short HEAD) [--workloads breaktime cornell] [--tag r2]

For every workload it captures, with `ncu --set full --clock-control none --import-source on`, EVERY launch of the
extend kernel (nearest-hit trace), the shadow-connect kernel (any-hit trace) and the shade kernel of one short, one-wave
render (tools/prof_run.py, CUDA graphs off so every launch is visible), reads the reports back with `ncu -i ... --page
raw --csv`, and writes

    gpurun_out/profiles/kernel_profiles.json  (copy to profiles/)
                                                per workload and kernel: duration, warp instructions (per ray / per hit),
                                         lanes per instruction, issue-slot utilisation, pipe utilisation, L1 / L2 hit
                                         rates, DRAM bytes (per ray / per hit) — with the git hash of the capture
    gpurun_out/profiles/<tag>_<kernel>_<workload>_ncu_summary.txt   the tracked metrics + the heaviest SASS basic blocks

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
PROFILES = os.path.join(OUT, "profiles")  # only gpurun_out/ travels back from the box: copy these into profiles/ afterwards

KERNELS = {
    # name: (ncu -k regex, kernel-name filter, unit of work)
    "extend": ("wf_trace", "<1,", "rays"),   # nearest-hit launches of the trace kernel: wf_trace_kernel<1, ...>
    "shadow": ("wf_trace", "<0,", "rays"),   # any-hit launches
    "shade": ("wf_shade", "wf_shade", "hits"),
}


def run_capture(workload, spp, regex, tag):
    """One render under ncu, every launch matching `regex` of its single wave captured.  Returns the report path and
    the per-bounce work counts parsed from the render's queue log."""
    rep = os.path.join(OUT, f"{tag}_{regex}_{workload}")
    env = dict(os.environ, RPT_GRAPHS="0", RPT_LOG_QUEUES="1")
    cmd = ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{regex}", "-c", "16", "-f", "-o", rep,
           sys.executable, os.path.join(REPO, "tools", "prof_run.py"), workload, str(spp)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if p.returncode != 0:
        raise SystemExit(f"ncu failed for {regex}/{workload}:\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
    # RPT_LOG_QUEUES: "[rpt] bounce b: extend traced N rays -> H hits, M misses; S shadow rays; P paths go on"
    work = {"extend": [], "shadow": [], "shade": []}
    for m in re.finditer(r"bounce (\d+): extend traced (\d+) rays -> (\d+) hits, (\d+) misses; (\d+) shadow rays", p.stderr):
        work["extend"].append(int(m.group(2)))
        work["shade"].append(int(m.group(3)))
        work["shadow"].append(int(m.group(5)))
    if not work["extend"]:
        raise SystemExit("queue log not found in the render's stderr:\n" + p.stderr[-2000:])
    return rep + ".ncu-rep", work


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]

    def metric(row, name):
        c = head.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[c], 1.0)
        return float(row[c].replace(",", "")) * scale

    return [(r[head.index("Kernel Name")], (lambda name, r=r: metric(r, name))) for r in rows[2:]], out


def summarize(rep, kernel, workload, work, tag, git_hash):
    """Per-bounce figures of `kernel` and their sums over the wave (what one step does is a repetition of this wave)."""
    _regex, name_filter, unit_name = KERNELS[kernel]
    launches, csv_text = raw_rows(rep)
    mine = [(name, metric) for name, metric in launches if name_filter in name.replace(" ", "").replace("(bool)", "") or (kernel == "shade" and "wf_shade" in name)]
    counts = work[kernel]  # the k-th launch of a kernel belongs to bounce k (a bounce without work still launches; it is skipped below)
    n = min(len(mine), len(counts))
    per_bounce, tot = [], {"units": 0.0, "ms": 0.0, "inst": 0.0, "thread_inst": 0.0, "dram": 0.0, "issue_weighted": 0.0}
    for b in range(n):
        name, metric = mine[b]
        units = counts[b]
        if units == 0:
            continue
        ms, inst = metric("gpu__time_duration.sum"), metric("smsp__inst_executed.sum")
        lanes = metric("smsp__thread_inst_executed_per_inst_executed.ratio")
        dram = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
        issue = metric("smsp__issue_active.avg.pct_of_peak_sustained_active")
        per_bounce.append({"bounce": b, unit_name: units, "duration_ms": ms, f"warp_instructions_per_{unit_name[:-1]}": inst / units, "lanes_per_instruction": lanes,
                           "issue_active_pct_of_peak": issue, f"dram_bytes_per_{unit_name[:-1]}": dram / units,
                           "l1_hit_pct": metric("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": metric("lts__t_sector_hit_rate.pct"),
                           "alu_pipe_pct": metric("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                           "fma_pipe_pct": metric("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                           "long_scoreboard_per_issue": metric("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                           "registers_per_thread": metric("launch__registers_per_thread"), "warps_active_pct": metric("sm__warps_active.avg.pct_of_peak_sustained_active")})
        tot["units"] += units; tot["ms"] += ms; tot["inst"] += inst; tot["thread_inst"] += inst * lanes; tot["dram"] += dram; tot["issue_weighted"] += issue * ms
    one = unit_name[:-1]
    entry = {
        "kernel_name": mine[0][0] if mine else "", "launches": f"all {len(per_bounce)} bounces of one {workload} wave ({spp_of(workload)} spp)", unit_name: tot["units"],
        "duration_ms": tot["ms"], f"warp_instructions_per_{one}": tot["inst"] / tot["units"], "lanes_per_instruction": tot["thread_inst"] / tot["inst"],
        "issue_active_pct_of_peak": tot["issue_weighted"] / tot["ms"], f"dram_bytes_per_{one}": tot["dram"] / tot["units"],
        "l1_hit_pct": sum(b["l1_hit_pct"] * b["duration_ms"] for b in per_bounce) / tot["ms"],
        "l2_hit_pct": sum(b["l2_hit_pct"] * b["duration_ms"] for b in per_bounce) / tot["ms"],
        "per_bounce": per_bounce, "source": os.path.basename(rep) + " (ncu --set full --clock-control none)",
    }
    csv_path = rep.replace(".ncu-rep", ".csv")
    with open(csv_path, "w") as f:
        f.write(csv_text)
    text = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_summary.py"), csv_path], capture_output=True, text=True).stdout
    blocks = subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_blocks.py"), rep, {"shade": "wf_shade"}.get(kernel, "wf_trace"), "14"],
                            capture_output=True, text=True).stdout
    with open(os.path.join(PROFILES, f"{tag}_{kernel}_{workload}_ncu_summary.txt"), "w") as f:
        f.write(f"# {tag} {entry['kernel_name'][:100]}\n# {entry['launches']}; git {git_hash}; ncu --set full --clock-control none\n")
        f.write(f"# one column per captured launch of the report ({os.path.basename(rep)}), in launch order\n")
        f.write("\n".join(line[:260] for line in text.splitlines() if ".max." not in line and ".min." not in line and ".sum.p" not in line) + "\n")
        f.write("\n# heaviest SASS basic blocks (share of warp instructions | instructions x executions | lanes enabled | stall samples | opcode mix)\n")
        f.write("\n".join(line[:220] for line in blocks.splitlines()) + "\n")
    return entry


_SPP = {}


def spp_of(workload):
    return _SPP.get(workload, "?")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--git-hash", required=True)
    ap.add_argument("--workloads", nargs="+", default=["breaktime", "cornell"])
    ap.add_argument("--kernels", nargs="+", default=list(KERNELS))
    ap.add_argument("--spp", type=int, default=4)
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--keep-reports", action="store_true", help="keep the .ncu-rep files (tens of MB each; gpurun brings back 64 MiB at most)")
    args = ap.parse_args()
    os.makedirs(PROFILES, exist_ok=True)
    path = os.path.join(PROFILES, "kernel_profiles.json")
    committed = os.path.join(REPO, "profiles", "kernel_profiles.json")  # workloads not captured this time keep their entries
    table = json.load(open(committed)) if os.path.exists(committed) else {}
    table["git"] = args.git_hash
    table["captured"] = datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ")
    table["how"] = "python tools/ncu_profile.py (ncu --set full --clock-control none, every launch of one wave of tools/prof_run.py, graphs off; sums over the bounces)"
    for workload in args.workloads:
        table.setdefault(workload, {})
        _SPP[workload] = args.spp
        reports = {}
        for kernel in args.kernels:
            regex = KERNELS[kernel][0]
            if regex not in reports:
                reports[regex] = run_capture(workload, args.spp, regex, args.tag)
            rep, work = reports[regex]
            table[workload][kernel] = summarize(rep, kernel, workload, work, args.tag, args.git_hash)
            print(workload, kernel, json.dumps({k: v for k, v in table[workload][kernel].items() if k != "per_bounce"})[:400], flush=True)
        if not args.keep_reports:
            for rep, _ in reports.values():
                os.remove(rep)
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
