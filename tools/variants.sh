# Build-variant experiments (run on the GPU box): each variant is built, parity-checked on the one-hot tests, and timed.
set -e
cd $GRAFT_REPO_ROOT
for v in "-DML2048_TMA_CHUNK_BYTES=32768 -DML2048_TMA_STAGES=3" "-DML2048_TMA_CHUNK_BYTES=24576 -DML2048_TMA_STAGES=4" "-DML2048_TMA_CHUNK_BYTES=16384 -DML2048_TMA_STAGES=4" "-DML2048_TMA_CHUNK_BYTES=49152 -DML2048_TMA_STAGES=2"; do
  ML2048_NVCC_EXTRA="$v" python -m ml2048_b200.build >/dev/null 2>&1 || python -c "from ml2048_b200 import build; build.build(force=True)"
  echo "variant [$v]"
  timeout 600 python -m pytest tests/test_cuda_parity.py -q -x -k "onehot" 2>&1 | tail -1
  for oh in f32 bf16 u8; do
    python tools/profile_core.py --onehot $oh --steps 30 --burn-in 256 | sed 's/.*prepare/prepare/' | cut -c1-100
  done
done
