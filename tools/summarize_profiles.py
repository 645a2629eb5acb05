"""Turn the raw ncu outputs that tools/profile.sh leaves in gpurun_out/ into the summaries committed under profiles/.

    python tools/summarize_profiles.py <tag>      e.g. v5  ->  profiles/r01_launches_unet128_<tag>.csv,
        r01_gemm_launch_metrics_unet128_<tag>.csv, r01_ncu_full_gemm_tc2_<tag>.csv, r01_ncu_full_gn_<tag>.csv, r01_traffic.json
"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "vN"
note = sys.argv[2] if len(sys.argv) > 2 else ""


def rows_of(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"') or l.startswith("ID,")][0]
    return list(csv.DictReader(lines[start:]))


def to_ns(v, u):
    return v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}[u]


# ---- launch list ---------------------------------------------------------------------------------------------------
out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 520: DCB_CUDA_GRAPH=0 python bench.py "
       f"--images 1 --steps 1 --warmup 3 --no-cpu (tools/profile.sh launches); {note}",
       "ID,Kernel Name,Block Size,Grid Size,Metric Name,Metric Unit,Metric Value"]
agg, tot = collections.OrderedDict(), 0.0
for r in rows_of(os.path.join(G, "launches.csv")):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = to_ns(float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    out.append(f'{r["ID"]},{name},"{r["Block Size"]}","{r["Grid Size"]}",gpu__time_duration.sum,ns,{int(v)}')
    a = agg.setdefault(name.replace("void ", "").split("<")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
out.append("# share of device time per kernel:")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"# {k:40s} n={n:4d} ms={v / 1e6:8.3f} share={v / tot:.3f}")
    print(out[-1])
open(os.path.join(P, f"r01_launches_unet128_{tag}.csv"), "w").write("\n".join(out) + "\n")

# ---- per-launch GEMM metrics -------------------------------------------------------------------------------------------
per = collections.OrderedDict()
for r in rows_of(os.path.join(G, "gemm_traffic.csv")):
    d = per.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"].split("(")[0]})
    v, u, n = float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r["Metric Name"]
    if n == "gpu__time_duration.sum":
        d["us"] = to_ns(v, u) / 1e3
    elif n.startswith("dram__bytes"):
        d["rd" if "read" in n else "wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    elif "pipe_tensor" in n:
        d["tp"] = v
    elif "hit_rate" in n:
        d["l2"] = v
    elif "per_second" in n:
        d["ghz"] = v * {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1}.get(u, 1)
rows = list(per.values())
out = [f"# ncu --metrics (time, dram bytes, tensor pipe active, L2 hit, clock) for the {len(rows)} tcgen05 GEMM launches of one "
       f"unet-128 pass; tools/profile.sh traffic; {note}", "idx,kernel,duration_us,dram_read_MB,dram_write_MB,tensor_active_pct,l2_hit_pct,sm_ghz"]
for i, d in enumerate(rows):
    out.append(f'{i},{d["kernel"]},{d["us"]:.1f},{d["rd"] / 1e6:.1f},{d["wr"] / 1e6:.1f},{d["tp"]:.1f},{d["l2"]:.1f},{d["ghz"]:.2f}')
open(os.path.join(P, f"r01_gemm_launch_metrics_unet128_{tag}.csv"), "w").write("\n".join(out) + "\n")
tot_us = sum(d["us"] for d in rows)
tot_b = sum(d["rd"] + d["wr"] for d in rows)
tw = sum(d["tp"] * d["us"] for d in rows) / tot_us
dom = max(rows, key=lambda d: d["us"])
j = {"unet128": tot_b / len(rows),
     "unet128_detail": {
         "what": f"ncu per-launch metrics of all {len(rows)} tcgen05 GEMM launches of one classify pass (1 image x 100 timesteps x 2 "
                 f"classes, S=200 samples per launch sequence), tools/profile.sh traffic -> profiles/r01_gemm_launch_metrics_unet128_{tag}.csv",
         "launches": len(rows), "sum_duration_ms": tot_us / 1e3, "sum_dram_bytes": tot_b,
         "avg_dram_bytes_per_launch": tot_b / len(rows), "time_weighted_tensor_pipe_active_pct": tw,
         "dominant_launch": {
             "shape": "M=3276800 (200 x 128 x 128 px) N=128 K=2304 (3x3 conv over the GroupNorm'd 256-channel concat)",
             "duration_us": dom["us"], "dram_bytes_read": dom["rd"], "dram_bytes_write": dom["wr"],
             "algorithmic_bytes_read": 1678311424, "algorithmic_bytes_write": 838860800,
             "tflops": 2 * 3276800 * 128 * 2304 / dom["us"] / 1e6, "tensor_pipe_active_pct": dom["tp"],
             "l2_sector_hit_pct": dom["l2"], "sm_clock_ghz": dom["ghz"]}},
     "unet128_dram_bytes_per_eval": tot_b / 200,
     "note": "unet128 = average DRAM bytes per GEMM launch of the captured pass (200 evals); bench.py scales the per-eval "
             "figure to its own launch size"}
json.dump(j, open(os.path.join(P, "r01_traffic.json"), "w"), indent=1)
print(f"GEMM launches {len(rows)}: {tot_us / 1e3:.2f} ms, {tot_b / 1e9:.1f} GB DRAM, tensor pipe (time weighted) {tw:.1f} %; dominant",
      json.dumps(j["unet128_detail"]["dominant_launch"]))

# ---- --set full captures -------------------------------------------------------------------------------------------------
M = ("dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,"
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,lts__t_sector_hit_rate.pct,"
     "sm__cycles_elapsed.avg.per_second,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active,"
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,"
     "smsp__inst_executed.avg.per_cycle_active,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active")
for rep, name in (("prof_gemm.ncu-rep", "gemm_tc2"), ("prof_gn.ncu-rep", "gn")):
    src = os.path.join(G, rep)
    if os.path.exists(src):
        r = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv", "--metrics", M], capture_output=True, text=True)
        open(os.path.join(P, f"r01_ncu_full_{name}_{tag}.csv"), "w").write(r.stdout)
