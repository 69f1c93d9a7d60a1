set -x
# round-1 final evidence (third session): tests, smoke, bench line, ingest-kernel timings + DRAM bytes, 200k all-vs-all, launch list
(time timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_final.log 2>&1; tail -4 gpurun_out/pytest_gpu_final.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 200 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c_final.err; tail -3 gpurun_out/bench_r1c_final.err; tail -c 300 gpurun_out/bench_r1c.json
timeout 60 python scripts/ingest_prof.py > gpurun_out/ingest_plain.log 2>&1 && \
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_compress|k_byte_hist' -c 60 --csv --log-file gpurun_out/r1c_ingest_kernels.csv python scripts/ingest_prof.py > gpurun_out/ncu_ingest.log 2>&1
cat gpurun_out/ingest_plain.log
timeout 150 python bench.py --workload allvsall --samples 200000 > gpurun_out/ava200k_n1_r1c.json 2> gpurun_out/ava200k_n1_r1c.err; tail -c 400 gpurun_out/ava200k_n1_r1c.json
timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/plain_r1c.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_r1c.log 2>&1
ls -la gpurun_out | tail -8
