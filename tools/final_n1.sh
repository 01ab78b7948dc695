set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01f_bench_ref.json 2> gpurun_out/r01f_bench_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/pre.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_ll.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"afsk_front_kernel|slicer_segments|slicer_verify|guard_fixup|gather_write|ax25_gap_kernel" -s 12 -c 14 -f -o gpurun_out/prof_r01f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_r01f.ncu-rep
tail -c 600 gpurun_out/r01f_bench_n1.json
