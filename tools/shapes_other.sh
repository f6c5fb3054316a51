# tools/shapes_other.sh: shapes outside the BASELINE configs (tools/bench_shape.py), default kernel choice
python tools/bench_shape.py rgb 1920 1080 256 256 cubic 0.7 256
python tools/bench_shape.py rgb 3840 2160 1280 720 lanczos - 64
python tools/bench_shape.py rgba 3840 2160 1280 720 lanczos - 64
python tools/bench_shape.py rgba 3840 2160 1920 1080 cubic - 64
python tools/bench_shape.py grey 4096 4096 1024 1024 cubic - 64
python tools/bench_shape.py r16g16b16 3000 2000 750 500 mitchel - 32
python tools/bench_shape.py r16g16b16a16 4096 4096 1024 1024 lanczos - 16
python tools/bench_shape.py rgba 4000 3000 800 600 mitchel - 32
python tools/bench_shape.py greya 3840 2160 1000 562 lanczos - 64
python tools/bench_shape.py rgb 2000 1500 900 675 catmulrom - 64
python tools/bench_shape.py grey 1920 1080 256 256 cubic 0.7 256
python tools/bench_shape.py rgb 1920 1080 3840 2160 cubic - 32
python tools/bench_shape.py rgba 1920 1080 3840 2160 lanczos - 32
python tools/bench_shape.py grey 2048 2048 4096 4096 mitchel - 32
python tools/bench_shape.py r16g16b16 1024 1024 3072 3072 catmulrom - 16
python tools/bench_shape.py rgb 3000 500 800 1500 lanczos - 16
python tools/bench_shape.py r16g16b16a16 3745 931 703 1280 box 1.3 8
