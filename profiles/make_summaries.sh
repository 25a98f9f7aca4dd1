#!/bin/bash
# Turns the captures a `bash scratch/profile_run.sh` gpurun call brought back (gpurun_out/<tag>_*) into the tracked
# summaries of this directory.  usage: bash profiles/make_summaries.sh r01d r01
set -e
tag=${1:-r01d}; out=${2:-r01}
cd "$(dirname "$0")/.."
for k in integrate solver; do
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv > /tmp/${tag}_$k.csv 2>/dev/null
  python profiles/ncu_pick.py /tmp/${tag}_$k.csv > profiles/${out}_${k}_ncu_full.txt
done
grep -v '^==' gpurun_out/${tag}_launches.csv > profiles/${out}_launches.csv
cp gpurun_out/${tag}_bench.json profiles/${out}_bench_line.json
python - "$tag" "$out" <<'PY'
import csv, json, sys
tag, out = sys.argv[1:3]
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
      "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches "
                "(gpurun_out/%s_*.ncu-rep; summaries in profiles/%s_*_ncu_full.txt)" % (tag, out)}
json.dump(tr, open("profiles/traffic.json", "w"), indent=1)
print(tr)
PY
