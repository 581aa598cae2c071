python tools/tune.py --sizes 1024 --gen and3 --paths fused --decode --reps 3 2>&1 | tail -1 | cut -c1-200
python tools/tune.py --sizes 4096 --gen mixed --paths fused --decode --reps 2 2>&1 | tail -1 | cut -c1-200
python tools/tune.py --sizes 32,64,128 --gen uniform --paths ws --decode --reps 5 2>&1 | tail -3 | cut -c1-200
