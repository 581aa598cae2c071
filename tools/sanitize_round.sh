#!/bin/bash
# Runs ON THE GPU BOX: compute-sanitizer over the shipped kernels (both encode kernels with both role maps of the
# warp-specialised one, both decode variants incl. the cp.async ring, the encoders' carry into stored words, compaction work units, index, the host pipeline).
#   bash tools/sanitize_round.sh <tag>
set -u
tag=${1:-rX}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
SEL='both_encode_kernels or both_decode_kernels or encode_long_underflow or decode_ragged_sizes or work_unit_sweep or index_equals_chain_walk or host_path_multi_chunk'
for tool in racecheck synccheck memcheck; do
  ( timeout 1500 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" 2>&1 | tail -6 ) \
      > gpurun_out/${tag}_sanitizer_$tool.txt 2>&1
  tail -n 3 gpurun_out/${tag}_sanitizer_$tool.txt
done
