# Scratch driver for core-only kernel experiments on the GPU box: bash tools/core_lab.sh <tag> [ncu|ncu2|variants|pdl]
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
if [ "$2" = "pdl" ]; then
for pdl in 0 1; do for g in 4194304 16777216; do echo "== ML2048_PDL=$pdl games=$g"; ML2048_PDL=$pdl python tools/profile_core.py --games $g --fused --steps 50 --burn-in 256; ML2048_PDL=$pdl python tools/profile_core.py --games $g --fused --rng philox --steps 50 --burn-in 256; done; done
ML2048_PDL=0 python tools/sweep.py --quick --out gpurun_out/sweep_pdl0.json > /dev/null 2>&1; python tools/sweep.py --quick --out gpurun_out/sweep_pdl1.json > /dev/null 2>&1
python - <<PY
import json
for f in ("gpurun_out/sweep_pdl0.json","gpurun_out/sweep_pdl1.json"):
    d=json.load(open(f)); print(f)
    for p in d["points"]:
        if p["rng"]=="replay": print("  ",p["config"], p["games"], p["onehot"], {k:round(v["us_per_step"],1) for k,v in p.items() if isinstance(v,dict)})
PY
exit 0
fi
run4
if [ "$2" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:"step_.*kernel" -s 262 -c 1 -o gpurun_out/prof_core_fused_$tag -f python tools/profile_core.py --steps 6 --burn-in 256 --fused > gpurun_out/ncu_core_fused_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_core_fused_$tag.log
fi
