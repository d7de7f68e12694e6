#!/usr/bin/env python
"""Summaries of Nsight Compute output for profiles/ (run in the build container; ncu reads the
.ncu-rep files brought back from the GPU box).

    python tools/ncu_summary.py rep  gpurun_out/X.ncu-rep  > profiles/X_ncu_full_summary.txt
    python tools/ncu_summary.py list gpurun_out/X_launches.csv > profiles/X_ncu_launches_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "gpc__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_tc.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fma.sum",
    # the SM <-> L2 path (what the stash, the weight ring and the dW reductions of the MLP kernels load)
    "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none: {path}")
    print("# per launch; dram bytes = measured traffic; times are under the profiler (serialised, cold cache)")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        print(f"\n== {name}")
        for k in KEYS:
            for h in hdr:
                if h == k or h.endswith("." + k):
                    print(f"  {k:78s} {r[idx[h]]:>18s} {units[idx[h]]}")
                    break
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
        tot += ms
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {path}")
    print(f"# launches captured: {sum(a[1] for a in agg.values())}, total {tot:.1f} ms "
          "(cold-cache, serialised: compare SHARES with bench.py's event-timed shares)")
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"{ms:12.3f} ms {n:5d}x {100 * ms / tot:6.2f}%  {name[:90]}")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
