set -x
python bench.py > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -c 600 gpurun_out/bench_r1_a.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_compare -c 20 -o gpurun_out/prof_r1_compare python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
