# usage: bash scripts/multi_gpu_batch8.sh N tag     (run under gpurun --gpus N): correctness, the bench line, c4, collectives
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
timeout 200 $TR --master-port 29721 tests/multigpu_check.py 2> gpurun_out/${TAG}_check.err | grep "^{" > gpurun_out/${TAG}_check.json; echo check rc=$?
timeout 400 $TR --master-port 29722 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench.err | grep "^{" > gpurun_out/${TAG}_bench.json; echo bench rc=$?
timeout 300 $TR --master-port 29723 scripts/run_c4.py 2> gpurun_out/${TAG}_c4.err | grep "^{" > gpurun_out/${TAG}_c4.json; echo c4 rc=$?
timeout 120 $TR --master-port 29724 scripts/time_comm.py 2> /dev/null > gpurun_out/${TAG}_time_comm.json; echo comm rc=$?
ERA5SVD_COMM=nccl timeout 200 $TR --master-port 29725 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-north-star 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_c2_nccl.json; echo bench nccl rc=$?
ERA5SVD_COMM=peer timeout 200 $TR --master-port 29726 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-north-star 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_c2_peer.json; echo bench peer rc=$?
