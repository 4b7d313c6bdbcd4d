# Multi-GPU evidence in ONE gpurun call (N x box time is charged):   bash scripts/multi_gpu_batch.sh N TAG [steps...]
#   steps (default: all): check comm bench c4 nccl peer pcie ref
# Outputs: gpurun_out/TAG_*.json (copied to profiles/ by hand once read).
N=$1; TAG=$2; shift 2
STEPS=${*:-check comm bench c4 nccl peer pcie ref}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
set -x
for s in $STEPS; do
  case $s in
    check) timeout 300 $TR --master-port 29721 tests/multigpu_check.py 2> gpurun_out/${TAG}_check.err | grep "^{" > gpurun_out/${TAG}_check.json ;;
    comm)  timeout 200 $TR --master-port 29722 scripts/time_comm.py 2> /dev/null > gpurun_out/${TAG}_time_comm.json ;;
    bench) timeout 400 $TR --master-port 29723 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/${TAG}_bench.err | grep "^{" > gpurun_out/${TAG}_bench.json ;;
    c4)    timeout 300 $TR --master-port 29724 scripts/run_c4.py 2> gpurun_out/${TAG}_c4.err | grep "^{" > gpurun_out/${TAG}_c4.json ;;
    nccl)  ERA5SVD_COMM=nccl timeout 200 $TR --master-port 29725 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-north-star 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_c2_nccl.json ;;
    peer)  ERA5SVD_COMM=peer timeout 200 $TR --master-port 29726 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-north-star 2> /dev/null | grep "^{" > gpurun_out/${TAG}_bench_c2_peer.json ;;
    pcie)  timeout 200 $TR --master-port 29727 scripts/pcie_probe_multi.py 2> gpurun_out/${TAG}_pcie_topology.txt | grep "^{" > gpurun_out/${TAG}_pcie.json ;;
    ref)   timeout 200 $TR --master-port 29728 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2> /dev/null | grep "^{" > gpurun_out/${TAG}_ref.json ;;
  esac
  echo "$s rc=$?"
done
