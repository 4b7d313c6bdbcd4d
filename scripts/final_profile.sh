# Round-end evidence on ONE GPU: tests, the bench line, then (only after both exited 0 without a profiler) the ncu launch
# list of a short bench run and one full capture of the tall kernels.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_tests.log 2>&1; echo tests rc=$?; tail -2 gpurun_out/r02f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r02f_smoke.log
python bench.py > gpurun_out/r02f_bench1.json 2> gpurun_out/r02f_bench1.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_ref1.json 2> /dev/null; echo ref rc=$?
python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-north-star > gpurun_out/r02f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-north-star > gpurun_out/r02f_ncu_list.log 2>&1; echo list rc=$?
ncu --set full --clock-control none --import-source on -k regex:"sketch_x1_kernel|project_tc_kernel|project_tc2_kernel|sketch_tc2_kernel|fused_build_tma" -s 10 -c 10 -f -o gpurun_out/r02f_prof python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-north-star > gpurun_out/r02f_ncu_full.log 2>&1; echo full rc=$?
