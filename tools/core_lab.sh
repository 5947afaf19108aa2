set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for i in 1 2; do
python tools/profile_core.py --fused --steps 30 --burn-in 256
python tools/profile_core.py --actions given --steps 30 --burn-in 256
done
python tools/profile_core.py --fused --rng philox --steps 30 --burn-in 256
python tools/profile_core.py --steps 30 --burn-in 256
