#!/bin/bash
# Turns the captures of `bash profiles/profile_run_r02.sh` (gpurun_out/<tag>_*) into the tracked round-2 summaries.
# usage: bash profiles/make_summaries_r02.sh r02f
set -e
tag=${1:-r02j}; out=r02
cd "$(dirname "$0")/.."
for k in integrate solver knn dense; do
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv > /tmp/${tag}_$k.csv 2>/dev/null
  python profiles/ncu_pick.py /tmp/${tag}_$k.csv > profiles/${out}_${k}_ncu_full.txt
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${tag}_${k}_src.csv 2>/dev/null
  python profiles/ncu_lines.py /tmp/${tag}_${k}_src.csv 30 > profiles/${out}_${k}_hot_lines.txt
done
grep -v '^==' gpurun_out/${tag}_launches.csv > profiles/${out}_launches.csv
cp gpurun_out/${tag}_bench.json profiles/${out}_bench_line.json
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
def per_launch(path, pattern):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    k, r, w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(x[r]) * scale[units[r]] + float(x[w]) * scale[units[w]] for x in rows[2:] if pattern in x[k]]
    return sum(vals) / len(vals) if vals else None
tr = {"integrate_dram_bytes_per_launch": per_launch("/tmp/%s_integrate.csv" % tag, "integrate_kernel<2>"),
      "solver_dram_bytes_per_launch": per_launch("/tmp/%s_solver.csv" % tag, "k_solve_persistent"),
      "dense_integrate_dram_bytes_per_launch": per_launch("/tmp/%s_dense.csv" % tag, "integrate_kernel<2>"),
      "knn_dram_bytes_per_launch": per_launch("/tmp/%s_knn.csv" % tag, "points_grid"),
      "source": "ncu --set full --clock-control none of `python bench.py` at N = 1 (sequential schedule), dram__bytes_read.sum + "
                "dram__bytes_write.sum, mean over the captured steady-state launches (profiles/profile_run_r02.sh; summaries in "
                "profiles/r02_*_ncu_full.txt)"}
json.dump(tr, open("profiles/traffic_r02.json", "w"), indent=1)
print(tr)
PY
# kernel shares of the step from the steady-state launch list
python - <<'PY'
import csv, collections
r = list(csv.reader(open("profiles/r02_launches.csv")))
h = r[0]; k = h.index("Kernel Name"); v = h.index("Metric Value")
d = collections.defaultdict(list)
for x in r[1:]:
    d[x[k].split("(")[0].replace("<unnamed>::", "").replace("void ", "")].append(float(x[v]))
tot = sum(sum(t) for t in d.values())
with open("profiles/r02_launch_shares.txt", "w") as f:
    f.write("steady-state launch list of `python bench.py --no-overlap` (ncu gpu__time_duration.sum, cold-cache, serialised): share of the summed kernel time\n")
    for n, t in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-48s launches %3d  mean %8.1f us  share %5.1f %%\n" % (n[:48], len(t), sum(t) / len(t) / 1e3, 100 * sum(t) / tot))
print(open("profiles/r02_launch_shares.txt").read())
PY
