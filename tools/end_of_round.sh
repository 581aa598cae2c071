#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): everything profiles/README.md cites for the end-of-round state.
#   bash tools/end_of_round.sh <tag>
set -u
tag=${1:-r1_end}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/${tag}_pytest.log 2>&1
bash tools/profile_round.sh $tag u64m > /dev/null 2>&1
bash tools/profile_round.sh $tag s1g > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/${tag}_bench_u64m.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --workload s1g --no-cpu-baseline > gpurun_out/${tag}_bench_s1g.json 2>> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --workload m2g --no-cpu-baseline > gpurun_out/${tag}_bench_m2g.json 2>> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err
# BASELINE config 5: packet-size sweep and work-unit sweep on 4 GiB (decode-heavy: 1 encode, steps decodes)
for p in 4096 8192 12288 16112; do
  timeout 600 python bench.py --workload m4g --packet $p --steps 5 --no-cpu-baseline \
      > gpurun_out/${tag}_sweep_m4g_packet$p.json 2>> gpurun_out/${tag}_bench.err
done
{
  timeout 300 python tools/tune.py --sizes 4096 --gen mixed --paths fused --tiles 4,8,16,32,64,128 --reps 3
  timeout 300 python tools/tune.py --sizes 64 --gen uniform --paths ws --tiles 4,8,16,32,64,128 --reps 10
  timeout 300 python tools/tune.py --sizes 8,16,32,48,64,96,112,128,192,256,384 --paths ws,fused --reps 5
} > gpurun_out/${tag}_work_unit_and_size_sweep.jsonl 2>> gpurun_out/${tag}_bench.err
for tool in ${SANITIZERS:-memcheck racecheck synccheck}; do   # SANITIZERS="" skips them (racecheck alone takes ~5 min)
  ( timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -q -x \
      -k "both_encode_kernels or decode_ragged_sizes or work_unit_sweep" 2>&1 | tail -6 ) \
      > gpurun_out/${tag}_sanitizer_$tool.txt 2>&1
done
tail -2 gpurun_out/${tag}_pytest.log
python tools/bench_brief.py gpurun_out/${tag}_bench_u64m.json gpurun_out/${tag}_bench_s1g.json gpurun_out/${tag}_bench_m2g.json
tail -3 gpurun_out/${tag}_bench.err
for f in gpurun_out/${tag}_sanitizer_*.txt; do tail -n 2 $f; done
