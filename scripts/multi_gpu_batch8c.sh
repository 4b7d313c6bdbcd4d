N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
timeout 300 $TR --master-port 29781 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2> gpurun_out/${TAG}_bench.err | grep "^{" > gpurun_out/${TAG}_bench.json; echo bench rc=$?
timeout 200 $TR --master-port 29782 scripts/run_c4.py 2> gpurun_out/${TAG}_c4.err | grep "^{" > gpurun_out/${TAG}_c4.json; echo c4 rc=$?
