"""profiles/r2_kernel_ncu.json from an `ncu --set full` capture of one minibatch: per kernel name the averages over
its launches of DRAM bytes (read + write), tensor-pipe activity and duration.
  ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_kernel_table.py /tmp/raw.csv"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
need = {"dur_us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
        "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "l2_to_sm": "l1tex__m_xbar2l1tex_read_bytes.sum"}
units = rows[1]
out = {}
for r in data:
  name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "")
  e = out.setdefault(name, {"n": 0, "dur_us": 0.0, "dram": 0.0, "tensor": 0.0, "warps": 0.0, "l2_to_sm": 0.0})
  f = lambda k: float(r[col[need[k]]].replace(",", "")) if need[k] in col and r[col[need[k]]] not in ("", "n/a") else 0.0
  scale = lambda k: {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(units[col[need[k]]], 1.0) if need[k] in col else 1.0
  e["n"] += 1
  e["dur_us"] += f("dur_us") * {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[col[need["dur_us"]]], 1.0)
  e["dram"] += f("dram_rd") * scale("dram_rd") + f("dram_wr") * scale("dram_wr")
  e["tensor"] += f("tensor"); e["warps"] += f("warps"); e["l2_to_sm"] += f("l2_to_sm") * scale("l2_to_sm")
table = {k: {"launches": e["n"], "us_per_launch_ncu": e["dur_us"] / e["n"], "dram_bytes_per_launch": e["dram"] / e["n"],
             "tensor_pipe_pct": e["tensor"] / e["n"], "warps_active_pct": e["warps"] / e["n"],
             "l2_to_sm_bytes_per_launch": e["l2_to_sm"] / e["n"]} for k, e in out.items()}
json.dump(table, open("profiles/r2_kernel_ncu.json", "w"), indent=1, sort_keys=True)
for k, e in sorted(table.items(), key=lambda kv: -kv[1]["us_per_launch_ncu"] * kv[1]["launches"]):
  print("%-28s n=%3d %7.1f us  dram %9.0f B  tensor %5.1f %%  warps %5.1f %%" % (
    k, e["launches"], e["us_per_launch_ncu"], e["dram_bytes_per_launch"], e["tensor_pipe_pct"], e["warps_active_pct"]))
