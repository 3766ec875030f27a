#!/bin/bash
# GPU parity tests only.  Usage: tools/gpu_tests.sh <tag> [pytest args]
tag=${1:-t}; shift || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs --durations=15 -s "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
grep -E "passed|failed|error" gpurun_out/${tag}_pytest.log | tail -5
grep -E "^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | head -20
