#!/bin/bash
# Large-batch support (ring ADAM kernel in K passes) and the ring kernel at batch 64.
tag=${1:-s3i}
out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > $out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.txt
DEEPIMPUTE_B200_ADAM=ring timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_multinet_gpu.py -m gpu -q > $out/pytest_ring.txt 2>&1; echo "pytest exit $?" >> $out/pytest_ring.txt
DEEPIMPUTE_B200_ADAM=ring DEEPIMPUTE_B200_ADAM_STORE=tma timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "single_step or epochs" > $out/pytest_ring_tma.txt 2>&1; echo "pytest exit $?" >> $out/pytest_ring_tma.txt
grep -E "passed|failed|exit" $out/pytest_gpu.txt $out/pytest_ring.txt $out/pytest_ring_tma.txt; grep -E "^FAILED|^E  " $out/pytest_*.txt | head -30
