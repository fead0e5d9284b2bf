timeout 120 ./build/test_conv check 2>&1 | grep -E "FAIL|failing|rror"
for m in 64 32 16; do echo "swap_min=$m"; FASTDET_SWAP_MIN=$m python tools/layer_times.py --reps 5 2>&1 | awk 'NR>=2 && NR<=12 || /total/'; done
