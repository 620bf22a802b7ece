#!/bin/bash
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error|placement stats|nursery|rejected" gpurun_out/${tag}_pytest_gpu.log | cut -c1-600
tail -30 gpurun_out/${tag}_pytest_gpu.log | cut -c1-300
