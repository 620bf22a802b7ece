#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ah_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2ah_pytest_gpu.log | cut -c1-300
bash tools/gpu_runs/bench3.sh
