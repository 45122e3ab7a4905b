"""profiles/traffic.json: mean DRAM bytes (read+write) per launch of each hot kernel, from an `ncu --set full` report.
Usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep   (or the `ncu -i ... --page raw --csv` dump of one)"""
import collections
import csv
import io
import json
import subprocess
import sys

NAMES = {"k_linearize_tile": "linearize_pm", "k_linearize_pm": "linearize_pm", "k_linearize_cm": "linearize_cm", "k_schur_cm": "schur_cm",
         "k_point_tile<0>": "spmv_pm", "k_point_tile<1>": "backsub_cost", "k_point_pass<0>": "spmv_pm", "k_point_pass<1>": "backsub_cost",
         "k_spmv_cm": "spmv_cm"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

raw = (open(sys.argv[1]).read() if sys.argv[1].endswith(".csv") else subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
acc = collections.defaultdict(list)
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("glba::", "")
    for k, v in NAMES.items():
        if name.startswith(k.split("<")[0]) and (("<" not in k) or k in r[ki].replace("glba::", "")):
            acc[v].append(float(r[ri]) * UNIT[units[ri]] + float(r[wi]) * UNIT[units[wi]])
            break
out = {k: sum(v) / len(v) for k, v in acc.items()}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(out)
