#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nsfhifigan or plumbing or hifigan_v1" 2>&1 | tail -8 | tee gpurun_out/nsf_pytest.log
