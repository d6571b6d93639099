#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/c46_pytest.log 2>&1
tail -4 gpurun_out/c46_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
