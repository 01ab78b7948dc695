timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r01h_bench_n1.json 2> gpurun_out/r01h_bench_n1.err
tail -c 200 gpurun_out/r01h_bench_n1.json
