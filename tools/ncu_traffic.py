"""Record the extend kernel's measured DRAM traffic per ray from an `ncu --set full` report.

usage: python tools/ncu_traffic.py REPORT.ncu-rep WORKLOAD RAYS_IN_LAUNCH [launch index]
Writes/updates profiles/extend_traffic.json, which bench.py reads for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, workload, rays = sys.argv[1], sys.argv[2], float(sys.argv[3])
    index = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    nearest = [r for r in data if "wf_trace_kernel<1>" in r[head.index("Kernel Name")] or "wf_trace_kernel<(bool)1>" in r[head.index("Kernel Name")]]
    row = nearest[index]

    def metric(name):
        c = head.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[c], 1.0)
        return float(row[c]) * scale

    total = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
    path = os.path.join(REPO, "profiles", "extend_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[workload] = {"dram_bytes_per_ray": total / rays, "dram_bytes_in_profiled_launch": total, "rays_in_profiled_launch": rays,
                       "issue_active_pct_of_peak": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                       "lanes_per_instruction": metric("smsp__thread_inst_executed_per_inst_executed.ratio"),
                       "warp_instructions_per_ray": metric("smsp__inst_executed.sum") / rays,
                       "source": os.path.basename(rep) + f" (ncu --set full, wf_trace_kernel<true> launch {index})"}
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)
    print(workload, table[workload])


if __name__ == "__main__":
    main()
