set -x
# 1) plain bench; 2) ncu launch list of the same program (short run); 3) one ncu --set full capture of k_query
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_query -s 10 -c 6 -o gpurun_out/prof_r1_query python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
