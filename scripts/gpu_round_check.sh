# One GPU-box visit: parity tests, bench line, ncu launch list + full capture of the dominant kernels.  Usage (from the
# repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_round_check.sh <tag>'
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants --no-target > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants --no-target > gpurun_out/ncu_${TAG}_1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass2_kernel -c 3 -o gpurun_out/prof_pass2_$TAG -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants --no-target > gpurun_out/ncu_${TAG}_2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dist_tma_kernel -c 1 -o gpurun_out/prof_dist_$TAG -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants --no-target > gpurun_out/ncu_${TAG}_3.log 2>&1
python bench.py --workload cfg5s --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-variants --no-target > gpurun_out/plain5s_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pass2_kernel -c 4 -o gpurun_out/prof_pass2_q2_$TAG -f python bench.py --workload cfg5s --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants --no-target > gpurun_out/ncu_${TAG}_4.log 2>&1
python scripts/wide_case.py > gpurun_out/wide_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wide_fit_kernel -c 1 -o gpurun_out/prof_wide_$TAG -f python scripts/wide_case.py > gpurun_out/ncu_${TAG}_5.log 2>&1
python scripts/bench_kgrid.py > gpurun_out/kgrid_$TAG.json 2> gpurun_out/kgrid_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
if [ "$2" = "cfg5" ]; then
python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline --no-variants > gpurun_out/bench_cfg5_1gpu_$TAG.json 2> gpurun_out/bench_cfg5_1gpu_$TAG.err; echo "cfg5 rc=$?"
fi
echo done
