for g in 4 8; do
export PICHA_B200_DOWN_G=$g
echo "== G=$g"
python tools/bench_shape.py rgb 1920 1080 256 256 cubic 0.7 256
python tools/bench_shape.py rgb 3840 2160 1280 720 lanczos - 64
python tools/bench_shape.py grey 4096 4096 1024 1024 cubic - 64
python tools/bench_shape.py r16g16b16 3000 2000 750 500 mitchel - 32
python tools/bench_shape.py rgba 4000 3000 800 600 mitchel - 32
python tools/bench_shape.py greya 3840 2160 1000 562 lanczos - 64
python tools/bench_shape.py rgb 2000 1500 900 675 catmulrom - 64
python tools/bench_shape.py grey 1920 1080 256 256 cubic 0.7 256
done
