# usage: bash scripts/multi_gpu_batch.sh N tag     (run under gpurun --gpus N)
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
timeout 300 $TR --master-port 29721 tests/multigpu_check.py > gpurun_out/${TAG}_check.json 2> gpurun_out/${TAG}_check.err; echo check rc=$?
timeout 200 $TR --master-port 29722 scripts/time_comm.py > gpurun_out/${TAG}_time_comm.json 2> /dev/null; echo comm rc=$?
for c in peer nccl; do
  ERA5SVD_COMM=$c timeout 300 $TR --master-port 29723 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-north-star 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_c2_$c.json; echo bench $c rc=$?
done
timeout 300 $TR --master-port 29724 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-north-star --no-collectives 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_c2_nocoll.json
python - <<PY
import json
for c in ("peer","nccl","nocoll"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_c2_%s.json"%c).read()); print(c, d["ms_per_step"], d["value"], d["config"]["collectives"][:30], d["clocks"])
    except Exception as e: print(c, "failed", e)
PY
