for c in 2 4 8 12 16 32; do echo "ctas/sm=$c"; FASTDET_C0_CTAS=$c python tools/layer_times.py --reps 5 2>&1 | awk 'NR==2'; done
