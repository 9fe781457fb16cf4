"""Turns an `ncu --set full` report and the launch list of the same command into the files committed under profiles/:
  <prefix>_kernels_ncu_raw.csv   one row per captured launch, the metrics DESIGN.md quotes (from `--page raw --csv`)
  <prefix>_traffic.json          dram bytes per launch of k_accumulate and of the sort kernels (what bench.py's roofline.traffic reads)
  <prefix>_launch_summary.txt    per-kernel share of the step from the gpu__time_duration launch list
usage: python tools/ncu_summary.py gpurun_out/r02_prof.ncu-rep gpurun_out/r02_launches.csv profiles/r02 <log_n> <windows>"""
import csv, json, subprocess, sys, collections

rep, launches, prefix, log_n, windows = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct"]
keep = [k for k in keep if k in idx]
with open(prefix + "_kernels_ncu_raw.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[idx[k]] for k in keep])
    for r in rows[2:]:
        w.writerow([r[idx[k]] for k in keep])

def unit_scale(u):
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)

entries = float(1 << log_n) * windows
traffic = {}
sort = 0.0
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    b = sum(float(r[idx[m]]) * unit_scale(units[idx[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    if name.startswith("k_accumulate") and "k_accumulate" not in traffic:
        traffic["k_accumulate"] = {"bytes": b, "entries": entries}
    for k in ("k_scatter_staged_b", "k_bucket_hist_b", "k_bucket_scatter_staged_b", "void k_decompose_b", "k_decompose_b"):
        if name.startswith(k) and k not in traffic.setdefault("_seen", {}):
            traffic["_seen"][k] = b
            sort += b
            break
traffic.pop("_seen", None)
traffic["sort"] = {"bytes": sort, "entries": entries, "kernels": "k_decompose_b + k_scatter_staged_b + k_bucket_hist_b + k_bucket_scatter_staged_b"}
traffic["capture"] = {"log_n": log_n, "windows": windows, "report": rep.split("/")[-1]}
json.dump(traffic, open(prefix + "_traffic.json", "w"), indent=1)

# launch list: shares per kernel
tot = collections.defaultdict(float)
cnt = collections.Counter()
with open(launches) as f:
    lines = [l for l in f if l.startswith('"')]
rd = list(csv.reader(lines))
h = rd[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
for r in rd[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(r[ui], 1.0)
    name = r[ki].split("(")[0]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
with open(prefix + "_launch_summary.txt", "w") as f:
    f.write(f"{sum(cnt.values())} launches, {total:.3f} ms of device time (cold-cache, serialised: compare shares)\n")
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"{v:10.3f} ms  {100 * v / total:6.2f} %  x{cnt[name]:4d}  {name}\n")
print(open(prefix + "_launch_summary.txt").read())
print(json.dumps(traffic, indent=1))
