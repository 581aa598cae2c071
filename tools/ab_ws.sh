# encode_ws role maps: GPUAR_B200_WS_TUNE="warps,map_first,map_later"
for lib in "" pk1; do
 if [ -n "$lib" ]; then export GPUAR_B200_LIB=$PWD/gpuar_b200/libgpuar_b200_$lib.so; else unset GPUAR_B200_LIB; fi
 for tune in "5,FFF54201,FFF54201" "8,5F4FF210,5F4FF210" "8,5FF1F240,5FF1F240" "8,5F1FF240,5F1FF240" "6,FF5F4210,FF5F4210"; do
  export GPUAR_B200_WS_TUNE="$tune"; echo "lib=$lib tune=$tune"
  python tools/tune.py --sizes 32,64,128,192 --gen uniform --paths ws --reps 4 2>&1 | tail -4 | cut -c60-130
 done
done
