# band-count sweep on two of the other shapes (PICHA_B200_BANDS overrides the planner's choice)
for b in 0 2 3 4 6 8 12; do
  if [ $b = 0 ]; then unset PICHA_B200_BANDS; else export PICHA_B200_BANDS=$b; fi
  echo "bands $b"
  python tools/bench_shape.py rgb 2000 1500 900 675 catmulrom - 64
  python tools/bench_shape.py rgb 3840 2160 1280 720 lanczos - 64
  python tools/bench_shape.py grey 4096 4096 1024 1024 cubic - 64
done
unset PICHA_B200_BANDS
python tools/bench_shape.py greya 3840 2160 1000 562 lanczos - 64
