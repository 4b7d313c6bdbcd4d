# compute-sanitizer passes over the small end-to-end case and the one-CTA factor kernels (sanitizer target).
#   bash scripts/sanitize.sh            (on a GPU box; writes gpurun_out/sanitize_*.log)
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -4 gpurun_out/sanitize_memcheck_smoke.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -k "chol_inv or syevj or gemm_f64" > gpurun_out/sanitize_racecheck_small.log 2>&1
echo "racecheck small factors rc=$?"; tail -4 gpurun_out/sanitize_racecheck_small.log
