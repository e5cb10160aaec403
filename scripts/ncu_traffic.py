#!/usr/bin/env python
"""profiles/traffic.json from a full ncu capture of ONE batch's map kernels (k_map_thread x 2, k_seed_scan, k_map x 2):
DRAM bytes read + written per step and kernel, stamped with the hash of the kernel sources they were captured from
(bench.py drops the figure when the sources have changed since).  usage: ncu_traffic.py report.ncu-rep [workload]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rep = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "gencode_synth"
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
out = {}
launches = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    name = d["Kernel Name"].split("<")[0].split("(")[0].replace("void ", "").replace("psa::", "").strip()
    def val(key):
        v = float(d[key].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]
    b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    out[name] = out.get(name, 0) + b
    launches[name] = launches.get(name, 0) + 1
tj = {"%s:%s" % (workload, k): int(v) for k, v in out.items()}
tj["kernel_source_sha16"] = bench.kernel_source_sha16()
tj["launches_summed"] = launches
tj["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per step (all launches of the kernel in one batch), ncu --set full, capture %s" % os.path.basename(rep)
json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(tj, indent=1))
