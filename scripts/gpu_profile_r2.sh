set -x
# round-2 evidence on ONE B200: 1) GPU tests; 2) the bench line; 3) ncu launch list of a short bench run; 4) ncu --set full of
# k_scan (ceiling 250 and no ceiling) and of k_query (one-vs-all at ceiling 20).  Every ncu command runs only after the same
# command has exited 0 without ncu.
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; tail -3 gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --quick --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --quick --no-cpu > gpurun_out/ncu1.log 2>&1
python scripts/scan_prof.py > gpurun_out/r2_scan_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 38 -c 2 -o gpurun_out/r2_scan250 python scripts/scan_prof.py > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 100 -c 2 -o gpurun_out/r2_scaninf python scripts/scan_prof.py > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_query -s 4 -c 2 -o gpurun_out/r2_query python bench.py --steps 5 --warmup 3 --quick --no-cpu > gpurun_out/ncu4.log 2>&1
cat gpurun_out/r2_scan_prof.log
ls -la gpurun_out | tail -12
