for v in - encdup un8t un2t; do
  if [ "$v" = "-" ]; then unset GPUAR_B200_LIB; else export GPUAR_B200_LIB=$PWD/gpuar_b200/libgpuar_b200_$v.so; fi
  python tools/tune.py --sizes 1024 --gen and3 --paths fused --decode --reps 3 2>&1 | tail -1 | cut -c1-200
  python tools/tune.py --sizes 4096 --gen mixed --paths fused --decode --reps 2 2>&1 | tail -1 | cut -c1-200
done
