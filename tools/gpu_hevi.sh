#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "hevi" > gpurun_out/pytest_hevi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_hevi.log
tail -40 gpurun_out/pytest_hevi.log
