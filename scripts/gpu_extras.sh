# One-GPU extras of a round: captures of the kernels the main check does not cover, batched-k timing, reference arm.
# Usage: gpurun --timeout 900 -- 'bash scripts/gpu_extras.sh <tag>'   (every step is bounded by its own timeout)
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
B="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-variants --no-target"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pass3_kernel -c 1 -o gpurun_out/prof_pass3_$TAG -f python bench.py $B > gpurun_out/ncu_${TAG}_a.log 2>&1; echo "pass3 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dist_tma_kernel -c 1 -o gpurun_out/prof_dist_$TAG -f python bench.py $B > gpurun_out/ncu_${TAG}_b.log 2>&1; echo "dist rc=$?"
timeout 120 python bench.py --workload cfg5s --steps 2 --warmup 1 --no-cpu-baseline --no-variants --no-target > gpurun_out/bench_cfg5s_$TAG.json 2> gpurun_out/bench_cfg5s_$TAG.err; echo "cfg5s rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:pass2_kernel -c 1 -o gpurun_out/prof_pass2_q2_$TAG -f python bench.py --workload cfg5s $B > gpurun_out/ncu_${TAG}_c.log 2>&1; echo "pass2 q2 rc=$?"
timeout 120 python scripts/wide_case.py > gpurun_out/wide_$TAG.log 2>&1; echo "wide rc=$?"; tail -1 gpurun_out/wide_$TAG.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:wide_fit_kernel -c 1 -o gpurun_out/prof_wide_$TAG -f python scripts/wide_case.py > gpurun_out/ncu_${TAG}_d.log 2>&1; echo "wide ncu rc=$?"
timeout 120 python scripts/bench_kgrid.py > gpurun_out/kgrid_$TAG.json 2> gpurun_out/kgrid_$TAG.err; echo "kgrid rc=$?"; cat gpurun_out/kgrid_$TAG.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
echo done
