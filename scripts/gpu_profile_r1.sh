set -x
# round-1 evidence, final code: 1) plain bench; 2) ncu launch list of the same program (short run);
# 3) one ncu --set full capture of k_query (one-vs-all FILTER + STREAM launches); 4) one of an all-vs-all tile
python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -c 400 gpurun_out/bench_r1b.err
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_query -s 10 -c 6 -o gpurun_out/prof_r1b_query python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
python scripts/ava_prof.py > gpurun_out/ava_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_query -s 4 -c 2 -o gpurun_out/prof_r1b_ava python scripts/ava_prof.py > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out | tail -15
