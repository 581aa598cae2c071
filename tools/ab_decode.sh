# decode timing of the library in use (GPUAR_B200_LIB selects a variant): 64/128 MiB uniform, 1 GiB and3
set -e
python tools/tune.py --sizes 64,128 --gen uniform --decode --reps 5
python tools/tune.py --sizes 1024 --gen and3 --decode --reps 3
