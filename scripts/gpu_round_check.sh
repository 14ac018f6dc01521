cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_r1d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r1d.log
tail -5 gpurun_out/pytest_r1d.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg4_r1d.json 2> gpurun_out/bench_cfg4_r1d.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/plain_d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_d1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass2_kernel -c 3 -o gpurun_out/prof_pass2_r1d -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_d2.log 2>&1
echo done
