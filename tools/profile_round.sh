#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): launch list + `--set full` captures of the kernels of one bench step.
#   bash tools/profile_round.sh <tag> [workload]
set -u
tag=${1:-rX}; wl=${2:-u64m}
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_${wl}_launches.csv \
    python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/${tag}_${wl}_ncu_a.log 2>&1
# the four kernels that carry the bytes, first timed step (12 matching launches belong to the warm-up)
ncu --set full --clock-control none --import-source on \
    -k regex:"encode_ws_kernel|encode_kernel|decode_kernel|compact_kernel|index_mark_kernel" -s 12 -c 4 \
    -o gpurun_out/${tag}_${wl}_full python bench.py --steps 1 --warmup 3 --workload $wl --no-cpu-baseline \
    > gpurun_out/${tag}_${wl}_ncu_b.log 2>&1
ls -la gpurun_out | grep ${tag}_${wl}
