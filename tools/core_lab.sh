# Scratch driver for core-only kernel experiments on the GPU box: bash tools/core_lab.sh <tag> [ncu]
cd $GRAFT_REPO_ROOT
tag=${1:-lab}
set -x
python tools/profile_core.py --fused --steps 30 --burn-in 256 || exit 1
python tools/profile_core.py --actions given --steps 30 --burn-in 256 || exit 1
python tools/profile_core.py --steps 30 --burn-in 256
python tools/profile_core.py --fused --rng philox --steps 30 --burn-in 256
if [ "$2" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel" -s 262 -c 1 -o gpurun_out/prof_core_fused_$tag -f python tools/profile_core.py --steps 6 --burn-in 256 --fused > gpurun_out/ncu_core_fused_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_core_fused_$tag.log
fi
