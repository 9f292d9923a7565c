set -x
python bench.py --steps 30 --warmup 5 > gpurun_out/r1l_bench_line.json 2> gpurun_out/r1l_bench.err
tail -c 600 gpurun_out/r1l_bench_line.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1l_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r1l_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_fusion_ -s 10 -c 5 -o gpurun_out/r1l_fusion python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r1l_ncu.log 2>&1
ls -la gpurun_out | tail -5
