#!/bin/bash
# Runs ON THE GPU BOX (under gpurun, one GPU): launch lists and `--set full` captures of the kernels of one step.
#   bash tools/profile_round.sh <tag>
set -u
tag=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
K='regex:encode_ws_kernel|encode_kernel|decode_kernel|compact_kernel|index_mark_kernel'
for wl in u64m s1g; do
  # every launch of a bench run with its device time (cold-cache, serialised: compare SHARES, not absolutes)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_${wl}_launches.csv \
      python bench.py --steps 2 --warmup 3 --workload $wl --sub '' --no-cpu-baseline > gpurun_out/${tag}_${wl}_ncu_a.log 2>&1
  # the kernels that carry the bytes, one profiled step
  ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" \
      -o gpurun_out/${tag}_${wl}_full -f python tools/ncu_target.py --workload $wl > gpurun_out/${tag}_${wl}_ncu_b.log 2>&1
done
# the headline workload (16 GiB): DRAM bytes and duration only -- one pass per kernel, nothing to save and restore
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k "$K" --csv --log-file gpurun_out/${tag}_m16g_dram.csv \
    python tools/ncu_target.py --workload m16g > gpurun_out/${tag}_m16g_ncu.log 2>&1
ls -la gpurun_out | grep ${tag}_ | head -20
