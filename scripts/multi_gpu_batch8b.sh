N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
timeout 200 $TR --master-port 29731 scripts/pcie_probe_multi.py 2> gpurun_out/${TAG}_pcie_topology.txt | grep "^{" > gpurun_out/${TAG}_pcie.json; echo pcie rc=$?
timeout 400 $TR --master-port 29732 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/${TAG}_bench.err | grep "^{" > gpurun_out/${TAG}_bench.json; echo bench rc=$?
timeout 200 $TR --master-port 29733 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2> /dev/null | grep "^{" > gpurun_out/${TAG}_ref.json; echo ref rc=$?
