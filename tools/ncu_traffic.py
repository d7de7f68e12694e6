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
           "k_march": "valid_samples", "k_density_bwd": "valid_samples",
           "k_appearance": "app_samples", "k_app_bwd": "app_samples_bwd"}
EXTRA = {"lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct_of_peak",
         "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
         "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = {"source": f"{rep} (ncu --set full --clock-control none, bench.py --rows 50)", "counts": counts, "kernels": {}}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
    if name not in unit_of:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[m]].replace(",", "")) * scale[units[ix[m]]]
    n = counts[unit_of[name]]
    e = {"unit": unit_of[name], "units_in_capture": n, "dram_bytes_in_capture": tot, "dram_bytes_per_unit": tot / n}
    def num(m):
        return float(r[ix[m]].replace(",", ""))
    if "lts__t_sectors.sum" in ix:      # lts__t_bytes = 32 B x sectors through the L2 tag stage
        lts = 32.0 * num("lts__t_sectors.sum")
        e.update(lts_bytes_in_capture=lts, lts_bytes_per_unit=lts / n)
        if "lts__t_sectors.sum.peak_sustained" in ix and "lts__cycles_elapsed.avg.per_second" in ix:
            hz = num("lts__cycles_elapsed.avg.per_second") * {"hz": 1, "Khz": 1e3, "Mhz": 1e6, "Ghz": 1e9}.get(
                units[ix["lts__cycles_elapsed.avg.per_second"]], 1)
            e["l2_peak_gbs"] = 32.0 * num("lts__t_sectors.sum.peak_sustained") * hz / 1e9
    for m, key in EXTRA.items():
        m = m if m in ix else next((h for h in hdr if h.endswith("." + m)), m)
        if m in ix:
            try:
                e[key] = float(r[ix[m]].replace(",", ""))
            except ValueError:
                pass
    res["kernels"][name] = e
print(json.dumps(res, indent=1))
