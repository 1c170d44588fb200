#!/bin/sh
# A/B timing of tuning builds: tools/ab_levels.sh <rounds> <lib or -> ...   (each round runs tools/ln_levels.py once per library)
rounds=$1; shift
for i in $(seq $rounds); do
  for lib in "$@"; do
    if [ "$lib" = "-" ]; then env -u MLE_CUDA_LIB python tools/ln_levels.py; else MLE_CUDA_LIB=$lib python tools/ln_levels.py; fi
  done
done
