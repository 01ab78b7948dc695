timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 400 python tools/verify_hour.py 2>&1 | tail -2 | tee gpurun_out/r01g_verify_hour.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r01g_bench_n1.json 2> gpurun_out/r01g_bench_n1.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/pre.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_ll.log 2>&1
tail -c 300 gpurun_out/r01g_bench_n1.json
