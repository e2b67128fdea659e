# timelines of the halo conv kernel on the smallest layers (CTA 0's clock64 stamps)
for shape in "fwd 128 128 3 8 8 1 1" "fwd 128 128 3 8 8 1 8" "dgrad 128 128 3 8 8 1 8" "fwd 132 128 3 16 16 1 8" "fwd 64 64 3 32 32 1 8"; do
  echo "== $shape"
  TC2_TIMELINE=1 python scripts/conv_probe.py $shape 20 2>&1 | grep -v "^$" | tail -8
done
