#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/c21_pytest.log 2>&1
tail -5 gpurun_out/c21_pytest.log
