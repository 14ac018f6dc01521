# Multi-GPU visit.  Usage: gpurun --gpus N --timeout 1200 -- 'bash scripts/gpu_multi_check.sh <tag> <N> [cfg5] [inproc]'
# (SKIP_TESTS=1 skips the parity tests, NOBAL=1 adds the run without the balanced distance exchange)
cd $GRAFT_REPO_ROOT
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
fi
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 3 --warmup 3 --no-variants > gpurun_out/bench_cfg4_n${N}_$TAG.json 2> gpurun_out/bench_cfg4_n${N}_$TAG.err; echo "bench cfg4 rc=$?"
if [ -n "$NOBAL" ]; then
NPDRB_NO_BALANCED=1 run 29512 --steps 3 --warmup 3 --no-variants --no-e2e > gpurun_out/bench_cfg4_n${N}_${TAG}_nobal.json 2> gpurun_out/bench_cfg4_n${N}_${TAG}_nobal.err; echo "bench unbalanced rc=$?"
fi
for a in "$@"; do
if [ "$a" = "cfg5" ]; then
run 29513 --workload cfg5 --steps 2 --warmup 1 --no-variants > gpurun_out/bench_cfg5_n${N}_$TAG.json 2> gpurun_out/bench_cfg5_n${N}_$TAG.err; echo "cfg5 rc=$?"
fi
if [ "$a" = "inproc" ]; then
python scripts/bench_inprocess.py --gpus $N --workload cfg5 --steps 2 --warmup 1 > gpurun_out/inproc_cfg5_n${N}_$TAG.json 2> gpurun_out/inproc_cfg5_n${N}_$TAG.err; echo "inproc cfg5 rc=$?"
python scripts/bench_inprocess.py --gpus $N --workload cfg4 --steps 2 --warmup 1 > gpurun_out/inproc_cfg4_n${N}_$TAG.json 2> gpurun_out/inproc_cfg4_n${N}_$TAG.err; echo "inproc cfg4 rc=$?"
fi
done
echo done
