#!/usr/bin/env bash
# Halo-conv timing experiments: normal, no-epilogue (1), no-TMA (2), neither (3) for the heaviest layer shapes.
# usage: bash scripts/conv_modes.sh > gpurun_out/conv_modes.txt
for shape in "fwd 36 16 3 256 256" "dgrad 36 16 3 256 256" "fwd 68 32 3 128 128" "dgrad 68 32 3 128 128" "fwd 132 64 3 64 64" "fwd 16 16 1 256 256"; do
  for mode in 0 1 2 3; do
    echo "== $shape  dbgmode=$mode"
    MFVI_TC2_DBGMODE=$mode TC2_TIMELINE=$([ $mode = 0 ] && echo 1) python scripts/conv_probe.py $shape 1 8 20 2>&1 | head -12
  done
done
