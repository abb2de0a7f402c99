"""Turn ncu captures brought back in gpurun_out/ into the committed summaries under profiles/.

usage: python scripts/summarise_profile.py <gpurun_out/dir> <round-tag>
  launches_<wl>.csv                    -> profiles/<tag>_launches_<wl>.md   (per-kernel mean device time and share)
  full_<wl>_<kernel>.ncu-rep (ncu -i)  -> profiles/<tag>_full_<kernel>.md   (the metrics B200_PROFILING.md names)
  and profiles/traffic.json            (dram bytes per launch of the captured kernel, read by bench.py)
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

for path in glob.glob(os.path.join(src, "launches_*.csv")):
    wl = os.path.basename(path)[len("launches_"):-4]
    rows = list(csv.DictReader([l for l in open(path) if l.startswith('"')]))
    agg = collections.OrderedDict()
    for r in rows:
        key = (r["Kernel Name"].split("(")[0].split("::")[-1], r["Grid Size"], r["Block Size"])
        agg.setdefault(key, []).append(float(r["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(out, "{}_launches_{}.md".format(tag, wl)), "w") as f:
        f.write("# ncu launch list, workload {} ({} launches)\n\n".format(wl, len(rows)))
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` around `bench.py --workload {} --steps 1 "
                "--warmup 0 --epochs 1`; per-launch times are cold-cache and serialised: compare SHARES.\n\n".format(wl))
        f.write("| kernel | grid | block | launches | mean ns | share |\n|---|---|---|---|---|---|\n")
        for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `{}` | {} | {} | {} | {:.0f} | {:.3f} |\n".format(name, grid, block, len(v), sum(v) / len(v), sum(v) / tot))

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "dram__bytes.sum.per_second"]
traffic_path = os.path.join(out, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for rep in glob.glob(os.path.join(src, "full_*.ncu-rep")):
    base = os.path.basename(rep)[len("full_"):-len(".ncu-rep")]
    wl, kernel = base.split("_", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(out, "{}_full_{}_{}.md".format(tag, wl, kernel)), "w") as f:
        f.write("# ncu --set full --clock-control none, kernel regex `{}`, workload {}\n\n".format(kernel, wl))
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].split("::")[-1]
            f.write("## `{}` grid {} block {}\n\n| metric | value | unit |\n|---|---|---|\n".format(
                name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for w in WANT:
                if w in hdr:
                    f.write("| {} | {} | {} |\n".format(w, r[hdr.index(w)], units[hdr.index(w)]))
            f.write("\n")
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")]); ru = units[hdr.index("dram__bytes_read.sum")]
                wr = float(r[hdr.index("dram__bytes_write.sum")]); wu = units[hdr.index("dram__bytes_write.sum")]
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                grid = r[hdr.index("Grid Size")]
                traffic.setdefault(wl, {})["{} grid {}".format(name, grid)] = int(rd * scale[ru] + wr * scale[wu])
            except (ValueError, KeyError):
                pass
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
print("wrote", sorted(os.listdir(out)))
