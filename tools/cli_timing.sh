#!/bin/bash
# Runs ON THE GPU BOX: BASELINE configs 1 and 2 through the `gpuar` command line (file -> file), with the
# tool's own statistics block (same fields as the reference's main.cpp:172-182) and the wall time around it.
#   bash tools/cli_timing.sh <tag>
set -u
tag=${1:-r1_end}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out /dev/shm/gpuar_cli
d=/dev/shm/gpuar_cli
python - <<PY
import sys
sys.path.insert(0, ".")
from gpuar_b200 import datagen as D
D.uniform(0x64, 64 << 20).tofile("$d/random_64m.dat")
D.and3(2, 1 << 30).tofile("$d/and3_1g.dat")
PY
out=gpurun_out/${tag}_cli_timing.txt
echo "GPUs visible: $(nvidia-smi -L | wc -l); host threads: $(nproc)" > $out
export GPUAR_B200_TRACE=1       # start-up costs (context, page-locking) on stderr
export GPUAR_B200_FAST_EXIT=1   # end the process without the CUDA teardown (opt-in; the default is orderly)
run() {
  echo "\$ gpuar $*" >> $out
  local t0=$(date +%s%N)
  ./gpuar_b200/gpuar "$@" 2>&1 | grep -v "^Start\|%\.\." >> $out
  local t1=$(date +%s%N)
  echo "wall $(( (t1 - t0) / 1000000 )) ms" >> $out
}
# config 1: the reference's CPU-runnable case (single host thread)
run c --host --in=$d/random_64m.dat --out=$d/host.gip
run d --host --in=$d/host.gip --out=$d/host.out
# config 2: the same file on the device.  A fresh box pages the driver and the binaries in during the
# first GPU processes, so every device command runs four times; the last three are the steady state.
for rep in 1 2 3 4; do run c --in=$d/random_64m.dat --out=$d/dev.gip; done
for rep in 1 2 3 4; do run d --in=$d/dev.gip --out=$d/dev.out; done
# config 3 through the files
for rep in 1 2 3; do run c --in=$d/and3_1g.dat --out=$d/and3.gip; done
for rep in 1 2 3; do run d --in=$d/and3.gip --out=$d/and3.out; done
{
  echo "payloads identical (bytes >= 20): $(cmp -i 20 $d/host.gip $d/dev.gip > /dev/null && echo yes || echo NO)"
  md5sum $d/random_64m.dat $d/host.out $d/dev.out | awk '{print $1}' | sort -u | wc -l | sed 's/^/distinct md5 over input, host round trip, device round trip: /'
  echo "payload md5 (from byte 20): $(tail -c +21 $d/dev.gip | md5sum | cut -d' ' -f1)  (golden u64m: b8cb9f0c...)"
  cmp $d/and3_1g.dat $d/and3.out > /dev/null && echo "1 GiB round trip: ok" || echo "1 GiB round trip: FAILED"
} >> $out
rm -rf $d
cat $out
