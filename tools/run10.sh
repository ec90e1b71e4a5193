#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_cli.log
grep -E "passed|failed" gpurun_out/pytest_cli.log | tail -2; grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_cli.log | head -20
python -c "import __graft_entry__ as g; g.smoke()"
