#!/bin/bash
# sibling front-ends (CLAP, HeAR): parity tests and a throughput sample
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_siblings.py tests/test_logmel_gpu.py -m gpu -q -x > gpurun_out/pytest_siblings.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_siblings.log
timeout 300 python tools/siblings_sample.py > gpurun_out/siblings_sample.json 2> gpurun_out/siblings_sample.err
echo "sample rc=$?"; cat gpurun_out/siblings_sample.json; tail -3 gpurun_out/siblings_sample.err
