#!/bin/bash
# Runs HERE after tools/profile_round.sh came back: turns gpurun_out/<tag>_* into the tracked files under profiles/,
# and profiles/traffic.json (DRAM bytes per launch, tied to the hash of the kernel sources they were captured from).
set -eu
tag=${1:-rX}
cd "$(dirname "$0")/.."
for w in u64m s1g; do
  python tools/launch_summary.py gpurun_out/${tag}_${w}_launches.csv > profiles/${tag}_${w}_launches.txt
  cp gpurun_out/${tag}_${w}_launches.csv profiles/
  python tools/ncu_summary.py gpurun_out/${tag}_${w}_full.ncu-rep > profiles/${tag}_${w}_ncu_full_summary.txt
  ncu -i gpurun_out/${tag}_${w}_full.ncu-rep --page raw --csv > profiles/${tag}_${w}_ncu_raw.csv 2>/dev/null
done
cp gpurun_out/${tag}_m16g_dram.csv profiles/
python - "$tag" <<'PY'
import csv, json, sys
sys.path.insert(0, ".")
import bench
tag = sys.argv[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def key_of(name):
    return ("encode" if "encode" in name else "compact" if "compact" in name else
            "index_mark" if "index_mark" in name else "decode" if "decode" in name else None)
out = {}
for w in ("u64m", "s1g"):
    rows = list(csv.reader(open(f"profiles/{tag}_{w}_ncu_raw.csv")))
    h, units = rows[0], rows[1]
    ik, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    d = {}
    for r in rows[2:]:
        k = key_of(r[ik])
        if k:
            d[k] = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
    out[w] = d
# the per-metric CSV of the 16 GiB pass
d = {}
hdr = None
for r in csv.reader(open(f"profiles/{tag}_m16g_dram.csv")):
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    row = dict(zip(hdr, r))
    k = key_of(row["Kernel Name"])
    if k and row["Metric Name"].startswith("dram__bytes"):
        d[k] = d.get(k, 0.0) + float(row["Metric Value"].replace(",", "")) * scale[row["Metric Unit"]]
out["m16g"] = d
for w in out:
    out[w]["kernel_sources_sha"] = bench.kernel_source_hash()
    out[w]["capture"] = f"profiles/{tag}_{w}_" + ("dram.csv" if w == "m16g" else "ncu_raw.csv")
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out))
PY
