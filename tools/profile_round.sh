cd $GRAFT_REPO_ROOT
set -x
# 1. plain runs first (each must exit 0 without ncu)
python bench.py --steps 3 --warmup 3 --burn-in 256 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/plain_r01_g.log 2>&1 || exit 1
python tools/profile_core.py --steps 6 --burn-in 256 --actions given >> gpurun_out/plain_r01_g.log 2>&1 || exit 1
# 2. launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 100 --csv --log-file gpurun_out/launches_r01_g.csv python bench.py --steps 3 --warmup 3 --burn-in 256 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/ncu_launches_r01_g.log 2>&1
# 3. full capture of the fused step kernel and of the fused prepare kernel (steady state)
ncu --set full --clock-control none --import-source on -k regex:"step_kernel|prepare_fused" -s 516 -c 2 -o gpurun_out/prof_step_r01_g -f python bench.py --steps 3 --warmup 3 --burn-in 256 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/ncu_step_r01_g.log 2>&1
# 4. full capture of the core-only step kernel with given actions
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 264 -c 1 -o gpurun_out/prof_core_given_r01_g -f python tools/profile_core.py --steps 6 --burn-in 256 --actions given > gpurun_out/ncu_core_r01_g.log 2>&1
tail -n 2 gpurun_out/ncu_step_r01_g.log; tail -n 2 gpurun_out/ncu_core_r01_g.log
