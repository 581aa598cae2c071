#!/bin/bash
# Runs ON THE GPU BOX (gpurun --gpus N): bench lines at N ranks and the multi-GPU parity check.
#   bash tools/scale_run.sh <tag> <N> [workloads...]
set -u
tag=$1; n=$2; shift 2
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
port=29520
for wl in "${@:-u64m}"; do
  port=$((port + 1))
  timeout 900 bash -c "$(declare -f run); n=$n; run $port bench.py --gpus $n --steps 10 --warmup 3 --workload $wl --no-cpu-baseline" \
      > gpurun_out/${tag}_scale_${wl}_n$n.json 2> gpurun_out/${tag}_scale_${wl}_n$n.err
  python tools/bench_brief.py gpurun_out/${tag}_scale_${wl}_n$n.json
done
timeout 600 bash -c "$(declare -f run); n=$n; run 29540 tests/multigpu_check.py" > gpurun_out/${tag}_multigpu_check_n$n.txt 2>&1
tail -4 gpurun_out/${tag}_multigpu_check_n$n.txt
