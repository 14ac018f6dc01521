# 2-GPU visit for the pipelined upload: bash scripts/gpu_n2_check.sh <tag>   (gpurun --gpus 2)
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
timeout 300 python -m pytest tests -m gpu -x -q -k "slab_events or ready_flags or multi_driver_two" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
run 29511 --steps 3 --warmup 3 --no-variants --no-target > gpurun_out/bench_cfg4_n2_$TAG.json 2> gpurun_out/bench_cfg4_n2_$TAG.err; echo "rc=$?"
NPDRB_NO_PIPELINED_GATHER=1 run 29512 --steps 3 --warmup 3 --no-variants --no-target > gpurun_out/bench_cfg4_n2_${TAG}_nopipe.json 2> gpurun_out/bench_cfg4_n2_${TAG}_nopipe.err; echo "rc=$?"
tail -5 gpurun_out/bench_cfg4_n2_$TAG.err
TAG=$TAG python - <<'PY'
import json, os
tag = os.environ["TAG"]
for f in (f"gpurun_out/bench_cfg4_n2_{tag}.json", f"gpurun_out/bench_cfg4_n2_{tag}_nopipe.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "inproc", d["e2e_inproc"] and round(d["e2e_inproc"]["ms_per_step"], 2),
              d["multi_gpu_check"] and d["multi_gpu_check"]["pair_list_hash_equal"])
    except Exception as e:
        print(f, "failed", e)
PY
