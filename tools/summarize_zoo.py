"""gpurun_out/r02_zoo.ncu-rep + gpurun_out/zoo.json (tools/kernel_zoo.py) -> profiles/r02_ncu_kernel_zoo.csv: one row per
profiled launch (`ncu --set full`) with the metrics the north star asks for -- tensor-pipe activity for the contractions, achieved
DRAM throughput for the norm / elementwise / DWT kernels -- and, per case, algorithmic work / summed duration against the
measured peaks (MEASURED_PEAKS.json: burst bf16 TF/s for a kernel timed alone, HBM copy GB/s)."""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
# the report itself (hundreds of MB with --import-source) stays on the GPU box: it is exported there with
#   ncu -i gpurun_out/r02_zoo.ncu-rep --page raw --csv > gpurun_out/r02_zoo_raw.csv
cases = [json.loads(l) for l in open(os.path.join(G, "zoo.json")) if l.startswith("{")]
rows = list(csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(G, "r02_zoo_raw.csv"))))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))


def val(r, name, scale=None):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    return v * mult if scale else v


out = [["case", "kernel", "grid", "block", "duration_us", "dram_read_MB", "dram_write_MB", "dram_pct_of_peak", "tensor_pipe_pct",
        "xu_pipe_pct", "issue_slots_pct", "sm_ghz", "case_work", "case_achieved", "case_frac_of_measured_peak"]]
k = 0
for c in cases:
    grp = data[k:k + c["launches"]]
    k += c["launches"]
    tot_us = sum(val(r, "gpu__time_duration.sum", True) or 0.0 for r in grp)
    if c["unit"] == "flops":
        ach = c["work"] / (tot_us * 1e-6) / 1e12
        frac, au = ach / peaks["bf16_tflops"], "TF/s of burst bf16 peak %.0f" % peaks["bf16_tflops"]
    else:
        ach = c["work"] / (tot_us * 1e-6) / 1e9
        frac, au = ach / peaks["hbm_gbs"], "GB/s of HBM copy peak %.0f" % peaks["hbm_gbs"]
    for j, r in enumerate(grp):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("dcb::", "")
        out.append([c["case"] if j == 0 else "", name, r[col["Grid Size"]], r[col["Block Size"]],
                    "%.1f" % (val(r, "gpu__time_duration.sum", True) or 0), "%.1f" % (val(r, "dram__bytes_read.sum", True) or 0),
                    "%.1f" % (val(r, "dram__bytes_write.sum", True) or 0),
                    "%.1f" % (val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0),
                    "%.1f" % (val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0),
                    "%.1f" % (val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active") or 0),
                    "%.1f" % (val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0),
                    "%.2f" % ((val(r, "sm__cycles_elapsed.avg.per_second") or 0) / (1e9 if units[col["sm__cycles_elapsed.avg.per_second"]] == "hz" else 1)),
                    ("%.4g %s" % (c["work"], c["unit"])) if j == 0 else "", ("%.1f %s" % (ach, au)) if j == 0 else "",
                    ("%.3f" % frac) if j == 0 else ""])
assert k == len(data), (k, len(data))
with open(os.path.join(P, "r02_ncu_kernel_zoo.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --profile-from-start off -k regex:dcb python tools/kernel_zoo.py (one B200, "
            "cold-cache serialised launches; every kernel of libdcb200.so at a shape it runs at in the bench workloads)\n")
    csv.writer(f).writerows(out)
for r in out:
    print(",".join(r[:2] + r[4:5] + r[7:11] + r[13:]))
