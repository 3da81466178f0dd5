#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vits or generator_fp32 or nsfhifigan" 2>&1 | tail -12 | tee gpurun_out/vits_pytest.log
