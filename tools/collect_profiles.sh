#!/bin/bash
# Runs HERE after tools/end_of_round.sh came back: turns gpurun_out/<tag>_* into the tracked files under profiles/.
set -eu
tag=${1:-r1_end}
cd "$(dirname "$0")/.."
for w in u64m s1g; do
  python tools/launch_summary.py gpurun_out/${tag}_${w}_launches.csv > profiles/${tag}_${w}_launches.txt
  cp gpurun_out/${tag}_${w}_launches.csv profiles/
  python tools/ncu_summary.py gpurun_out/${tag}_${w}_full.ncu-rep > profiles/${tag}_${w}_ncu_full_summary.txt
  ncu -i gpurun_out/${tag}_${w}_full.ncu-rep --page raw --csv > profiles/${tag}_${w}_ncu_raw.csv 2>/dev/null
done
for f in gpurun_out/${tag}_bench_*.json gpurun_out/${tag}_sweep_*.json gpurun_out/${tag}_scale_*.json \
         gpurun_out/${tag}_work_unit_and_size_sweep.jsonl gpurun_out/${tag}_sanitizer_*.txt gpurun_out/${tag}_multigpu_check_n8.txt; do
  [ -e "$f" ] && case "$f" in *_full.txt) ;; *) cp "$f" profiles/ ;; esac
done
python - "$tag" <<'PY'
import csv, json, sys
tag = sys.argv[1]
out = {}
for w in ("u64m", "s1g"):
    rows = list(csv.reader(open(f"profiles/{tag}_{w}_ncu_raw.csv")))
    h, units = rows[0], rows[1]
    ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    d = {}
    for r in rows[2:]:
        name = r[ik]
        key = ("encode" if "encode" in name else "compact" if "compact" in name else
               "index_mark" if "index_mark" in name else "decode" if "decode" in name else None)
        if key:
            d[key] = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    out[w] = d
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
