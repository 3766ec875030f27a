#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -v -k "mini-alexnet or predict_classifier or demo_pipeline or validation_files or multi_gpu" -rs > $out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "PASSED|FAILED|ERROR|passed|failed|Abort" $out/r2g_pytest.log | tail -30
