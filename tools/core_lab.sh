# Scratch driver for core-only kernel experiments on the GPU box: bash tools/core_lab.sh <tag> [ncu|ncu2|variants]
cd $GRAFT_REPO_ROOT
tag=${1:-lab}
run4() {
python tools/profile_core.py --fused --steps 30 --burn-in 256 || exit 1
python tools/profile_core.py --actions given --steps 30 --burn-in 256 || exit 1
python tools/profile_core.py --steps 30 --burn-in 256
python tools/profile_core.py --fused --rng philox --steps 30 --burn-in 256
}
if [ "$2" = "variants" ]; then
for lib in gpurun_scratch_libs/*.so; do echo "== $lib"; ML2048_LIB=$PWD/$lib run4; done
echo "== default"; run4
exit 0
fi
run4
if [ "$2" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel" -s 262 -c 1 -o gpurun_out/prof_core_fused_$tag -f python tools/profile_core.py --steps 6 --burn-in 256 --fused > gpurun_out/ncu_core_fused_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_core_fused_$tag.log
fi
if [ "$2" = "ncu2" ]; then
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel" -s 262 -c 1 -o gpurun_out/prof_core_fused_$tag -f python tools/profile_core.py --steps 6 --burn-in 256 --fused > gpurun_out/ncu_core_fused_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel" -s 258 -c 1 -o gpurun_out/prof_core_random_$tag -f python tools/profile_core.py --steps 6 --burn-in 256 > gpurun_out/ncu_core_random_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_core_fused_$tag.log gpurun_out/ncu_core_random_$tag.log
fi
