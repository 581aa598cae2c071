# usage: tools/ab_variants.sh "<variant names>" <tune.py args...>   ("" = the product library)
names=$1; shift
for v in $names; do
  if [ "$v" = "-" ]; then unset GPUAR_B200_LIB; else export GPUAR_B200_LIB=$PWD/gpuar_b200/libgpuar_b200_$v.so; fi
  python tools/tune.py "$@" 2>&1 | tail -4
done
