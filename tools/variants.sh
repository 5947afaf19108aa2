set -e
cd $GRAFT_REPO_ROOT
for v in "-DML2048_STORE_DEFAULT" "-DML2048_STORE_DEFAULT -DML2048_STEP_THREADS=512" "-DML2048_STORE_DEFAULT -DML2048_STEP_THREADS=1024" "-DML2048_STORE_DEFAULT -DML2048_STEP_THREADS=768"; do
  ML2048_NVCC_EXTRA="$v" python -m ml2048_b200.build >/dev/null 2>&1
  echo "variant [$v]"
  for oh in f32 bf16 u8 none; do
    if [ $oh = none ]; then python tools/profile_core.py --steps 20 | sed 's/.*prepare/prepare/' | cut -c1-90; else
    python tools/profile_core.py --onehot $oh --steps 20 | sed 's/.*prepare/prepare/' | cut -c1-90; fi
  done
done
