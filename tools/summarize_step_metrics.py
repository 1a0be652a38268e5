"""profiles/r1_step_metrics.csv (ncu, one steady-state step) -> per-kernel summary + conv traffic json."""
import csv, collections, json, sys
src = sys.argv[1] if len(sys.argv) > 1 else "profiles/r1_step_metrics.csv"
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi, ui, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
per = collections.OrderedDict()
for r in rows:
    if r is hdr or not r[0].isdigit():
        continue
    d = per.setdefault(r[idi], {"name": r[ki].split("(")[0].replace("dynmm::<unnamed>::", "").replace("dynmm::", "")})
    v, u = float(r[vi].replace(",", "")), r[ui]
    if r[mi].startswith("dram__bytes") or r[mi].startswith("lts__t_bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if r[mi] == "gpu__time_duration.sum":
        v *= {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(u, 1e-9)
    d[r[mi]] = v
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"][:48], {"n": 0, "t": 0, "rd": 0, "wr": 0, "l2": 0, "tw": 0})
    t = d.get("gpu__time_duration.sum", 0)
    a["n"] += 1; a["t"] += t
    a["rd"] += d.get("dram__bytes_read.sum", 0); a["wr"] += d.get("dram__bytes_write.sum", 0)
    a["l2"] += d.get("lts__t_bytes.sum", 0)
    a["tw"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * t
tot = sum(a["t"] for a in agg.values())
out = [f"{'kernel':50s} {'n':>4s} {'us':>8s} {'share':>6s} {'DRAM rd MB':>10s} {'DRAM wr MB':>10s} {'L2 MB':>9s} {'DRAM GB/s':>9s} {'tensor%':>7s}"]
conv = {"n": 0, "t": 0, "rd": 0, "wr": 0, "l2": 0, "tw": 0}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    out.append(f"{k:50s} {a['n']:4d} {a['t']*1e6:8.1f} {100*a['t']/tot:5.1f}% {a['rd']/1e6:10.1f} {a['wr']/1e6:10.1f} "
               f"{a['l2']/1e6:9.1f} {(a['rd']+a['wr'])/a['t']/1e9:9.0f} {a['tw']/a['t'] if a['t'] else 0:7.1f}")
    if "conv_igemm" in k or "conv_chain" in k or "conv_pair" in k:
        for f in conv:
            conv[f] += a[f]
out.append(f"{'TOTAL (serialised, cold L2 per launch)':50s} {sum(a['n'] for a in agg.values()):4d} {tot*1e6:8.1f}")
out.append("")
out.append(f"tensor-core conv kernels (conv_igemm / conv_pair / conv_chain, all variants): {conv['n']} launches, {conv['t']*1e6:.1f} us ({100*conv['t']/tot:.1f}% of GPU time), "
           f"DRAM {conv['rd']/1e6:.1f} MB read + {conv['wr']/1e6:.1f} MB written = {(conv['rd']+conv['wr'])/conv['n']/1e6:.2f} MB per launch, "
           f"L2 traffic {conv['l2']/1e6:.0f} MB, time-weighted tensor pipe active {conv['tw']/conv['t']:.1f}%")
open(src.replace(".csv", "_summary.txt"), "w").write("\n".join(out) + "\n")
json.dump({"kernel": "conv_igemm_kernel + conv_pair_kernel + conv_chain_kernel", "launches_per_step": conv["n"], "dram_bytes_per_launch": (conv["rd"] + conv["wr"]) / conv["n"],
           "dram_read_bytes_per_step": conv["rd"], "dram_write_bytes_per_step": conv["wr"], "l2_bytes_per_step": conv["l2"],
           "tensor_pipe_active_pct_time_weighted": conv["tw"] / conv["t"], "kernel_time_s_per_step_serialized": conv["t"],
           "share_of_gpu_time": conv["t"] / tot,
           "source": src + " (ncu --metrics gpu__time_duration.sum,dram__bytes_*,lts__t_bytes,sm__pipe_tensor_cycles_active; "
                     "one steady-state step of the bench workload, eager launches, B=8 480x640, branches [0,4,0,0,4,4,4,0])"},
          open(src.replace("step_metrics.csv", "conv_traffic.json"), "w"), indent=1)
print("\n".join(out))
