# Convenience targets (the driver uses __graft_entry__.py, pytest and bench.py directly).
PY ?= python

build:            ## nvcc -gencode arch=compute_100a,code=sm_100a -> ml2048_b200/libml2048_b200.so, gcc -> oracle/liboracle.so
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## oracle vs golden fixtures, ABI, host-compiled device arithmetic, gloo sharding, bench contract
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## parity proper (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench:            ## one JSON line (see bench.py)
	$(PY) bench.py

golden:           ## regenerate tests/golden from the live reference (authoring container only)
	PYTHONDONTWRITEBYTECODE=1 $(PY) -m oracle.gen_golden

.PHONY: build test-cpu test-gpu bench golden
