#!/usr/bin/env python
"""DRAM traffic per unit of work of the top kernels, from one `ncu --set full` capture of
`bench.py --rows R` (tools/gpu_round.sh) -> profiles/ncu_traffic.json, which bench.py scales by the
units of its own run for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/X_prof.ncu-rep gpurun_out/X_ncu_full.log > profiles/ncu_traffic.json
"""
import csv
import io
import json
import re
import subprocess
import sys

rep, log = sys.argv[1], sys.argv[2]
counts = json.loads("{" + re.findall(r'"counts": \{([^}]*)\}', open(log).read())[-1] + "}")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
unit_of = {"k_advect_bwd_tc": "advected_samples_bwd", "k_sample_advect_tc": "advected_samples",
           "k_advect_bwd_h": "advected_samples_bwd", "k_sample_advect_h": "advected_samples",
           "k_march": "advected_samples", "k_density_bwd": "valid_samples",
           "k_appearance": "app_samples", "k_app_bwd": "app_samples_bwd"}
EXTRA = {"lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct_of_peak",
         "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
         "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {"source": f"{rep} (ncu --set full --clock-control none, bench.py --rows 50)", "counts": counts, "kernels": {}}
# one full step is captured (tools/gpu_round.sh): the wave kernels of the forward pass appear once per depth
# wave, so bytes are SUMMED per kernel over the step and divided by the step's units; percentages are averaged
acc = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
    if name not in unit_of:
        continue

    def num(m):
        return float(r[ix[m]].replace(",", ""))
    a = acc.setdefault(name, {"dram": 0.0, "lts": 0.0, "n": 0, "pct": {}})
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        a["dram"] += num(m) * scale[units[ix[m]]]
    if "lts__t_sectors.sum" in ix:
        a["lts"] += 32.0 * num("lts__t_sectors.sum")
    a["n"] += 1
    for m, key in EXTRA.items():
        m = m if m in ix else next((h for h in hdr if h.endswith("." + m)), m)
        if m in ix:
            try:
                a["pct"][key] = a["pct"].get(key, 0.0) + num(m)
            except ValueError:
                pass
for name, a in acc.items():
    n = counts[unit_of[name]]
    e = {"unit": unit_of[name], "units_in_capture": n, "launches_in_capture": a["n"], "dram_bytes_in_capture": a["dram"],
         "dram_bytes_per_unit": a["dram"] / n}
    if a["lts"]:
        e.update(lts_bytes_in_capture=a["lts"], lts_bytes_per_unit=a["lts"] / n)
    for key, v in a["pct"].items():
        e[key] = v / a["n"]
    res["kernels"][name] = e
print(json.dumps(res, indent=1))
