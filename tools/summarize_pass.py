"""ncu csv logs of tools/one_pass.py -> profiles/:
    python tools/summarize_pass.py launches <csv> <out name> "<note>"   every launch + share of device time per kernel
    python tools/summarize_pass.py gemm <csv> <out name> <evals> "<note>"   per-launch metrics of the tcgen05 GEMM launches
                                                                            (+ profiles/r02_traffic.json for bench.py)"""
import collections, csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def rows_of(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"') or l.startswith("ID,")][0]
    return list(csv.DictReader(lines[start:]))


def to_ns(v, u):
    return v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}[u]


mode, src, name = sys.argv[1:4]
if mode == "launches":
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off: {note}",
           "ID,Kernel Name,Block Size,Grid Size,Metric Name,Metric Unit,Metric Value"]
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows_of(src):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        kn = re.sub(r"\(.*", "", r["Kernel Name"])
        v = to_ns(float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
        out.append(f'{r["ID"]},{kn},"{r["Block Size"]}","{r["Grid Size"]}",gpu__time_duration.sum,ns,{int(v)}')
        a = agg.setdefault(kn.replace("void ", "").split("<")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out.append(f"# total device time {tot / 1e6:.3f} ms; share per kernel:")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"# {k:44s} n={n:4d} ms={v / 1e6:8.3f} share={v / tot:.3f}")
        print(out[-1])
    open(os.path.join(P, name), "w").write("\n".join(out) + "\n")
else:
    evals = int(sys.argv[4])
    note = sys.argv[5] if len(sys.argv) > 5 else ""
    per = collections.OrderedDict()
    for r in rows_of(src):
        d = per.setdefault(int(r["ID"]), {"kernel": re.sub(r"\(.*", "", r["Kernel Name"]).replace("dcb::", "")})
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
    out = [f"# ncu per-launch metrics of the {len(rows)} tcgen05 GEMM launches of one pass ({evals} evals): {note}",
           "idx,kernel,duration_us,dram_read_MB,dram_write_MB,tensor_active_pct,l2_hit_pct,sm_ghz"]
    for i, d in enumerate(rows):
        out.append(f'{i},{d["kernel"]},{d["us"]:.1f},{d["rd"] / 1e6:.1f},{d["wr"] / 1e6:.1f},{d["tp"]:.1f},{d["l2"]:.1f},{d["ghz"]:.2f}')
    tot_us = sum(d["us"] for d in rows)
    tot_b = sum(d["rd"] + d["wr"] for d in rows)
    tw = sum(d["tp"] * d["us"] for d in rows) / tot_us
    by = collections.OrderedDict()
    for d in rows:
        a = by.setdefault(d["kernel"], [0, 0.0, 0.0])
        a[0] += 1; a[1] += d["us"]; a[2] += d["tp"] * d["us"]
    out.append(f"# all: {tot_us / 1e3:.2f} ms, {tot_b / 1e9:.2f} GB DRAM, time-weighted tensor pipe active {tw:.1f} %")
    for k, (n, us, tpus) in by.items():
        out.append(f"# {k}: n={n} {us / 1e3:.2f} ms, tensor pipe {tpus / us:.1f} %")
    open(os.path.join(P, name), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-4:]))
    tj = os.path.join(P, "r02_traffic.json")
    j = json.load(open(tj)) if os.path.exists(tj) else {}
    wl = name.split("_")[-1].replace(".csv", "")
    j[wl + "_dram_bytes_per_eval"] = tot_b / evals
    j[wl + "_detail"] = {"launches": len(rows), "sum_duration_ms": tot_us / 1e3, "sum_dram_bytes": tot_b,
                         "time_weighted_tensor_pipe_active_pct": tw, "source": name, "evals": evals}
    json.dump(j, open(tj, "w"), indent=1)
