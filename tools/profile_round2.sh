# Round-2 profiling pass (run on the GPU box through gpurun): every command first runs plain and must exit 0, then under ncu.
#   bash tools/profile_round2.sh <tag>        e.g. r02_g
cd $GRAFT_REPO_ROOT
tag=${1:-r02}
B="python bench.py --steps 3 --warmup 3 --burn-in 256 --no-e2e --no-extras --no-cpu-baseline --no-shard-check"
set -x
# 1. plain runs
$B > gpurun_out/plain_$tag.log 2>&1 || exit 1
python tools/profile_core.py --steps 6 --burn-in 256 --fused >> gpurun_out/plain_$tag.log 2>&1 || exit 1
python tools/profile_core.py --steps 6 --burn-in 256 --actions given >> gpurun_out/plain_$tag.log 2>&1 || exit 1
# 2. launch list of the bench command (tail of the burn-in + warm-up + timed steps: steady state)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_.*kernel|prepare|autoreset" -s 500 -c 60 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_launches_$tag.log 2>&1
# 3. full capture: the fused fp32 step kernel (auto-reset inside) and the scan that precedes it, in the timed region
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel|autoreset" -s 262 -c 2 -o gpurun_out/prof_step_$tag -f $B > gpurun_out/ncu_step_$tag.log 2>&1
# 4. full capture: the core-only step kernel with the auto-reset fused in (random policy), and with given actions
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel" -s 262 -c 1 -o gpurun_out/prof_core_fused_$tag -f python tools/profile_core.py --steps 6 --burn-in 256 --fused > gpurun_out/ncu_core_fused_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel" -s 264 -c 1 -o gpurun_out/prof_core_given_$tag -f python tools/profile_core.py --steps 6 --burn-in 256 --actions given > gpurun_out/ncu_core_given_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_step_$tag.log gpurun_out/ncu_core_fused_$tag.log gpurun_out/ncu_core_given_$tag.log
cat gpurun_out/plain_$tag.log | tail -4
