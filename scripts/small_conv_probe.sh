# timelines of the halo conv kernel on the smallest layers (one CTA's clock64 stamps) + the verbose tile plan
for shape in "fwd 128 128 3 8 8 1 1" "fwd 128 128 3 8 8 1 8" "dgrad 128 128 3 8 8 1 1" "fwd 128 128 3 16 16 1 1" "fwd 64 64 3 32 32 1 1"; do
  for deep in 0 190; do
    echo "== $shape  MFVI_TC2_DEEP=$deep"
    MFVI_TC2_DEEP=$deep MFVI_TC2_VERBOSE=1 TC2_TIMELINE=1 python scripts/conv_probe.py $shape 20 2>&1 | grep -v "^$" | head -14
  done
done
