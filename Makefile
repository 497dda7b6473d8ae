# Top-level convenience targets (the driver uses __graft_entry__.build() / pytest / bench.py directly).
PY ?= python

build:            ## library + CLI (nvcc, sm_100a), oracle, host twin of the device sincosf
	$(PY) -c 'import __graft_entry__ as g; g.build(); print("build ok")'

test: build       ## CPU suite (oracle, planner, work decomposition, ABI, CLI argv, orbit, gloo slices, bench contract)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## parity suite on a B200
	$(PY) -m pytest tests -x -q -m gpu

bench: build      ## one JSON line (BASELINE metric)
	$(PY) bench.py

tune:             ## tuning harnesses (GPU box)
	$(MAKE) -C tools/tune all

clean:
	$(MAKE) -C doppler_b200/csrc clean
	$(MAKE) -C oracle clean
	$(MAKE) -C tests/native clean
	$(MAKE) -C tools/tune clean

.PHONY: build test test-gpu bench tune clean
