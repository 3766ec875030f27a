#!/bin/bash
out=gpurun_out; mkdir -p $out
for sel in "multi_gpu" "validation_files" "reference_detector_golden or do_nms or sparse_tree" "yolo9000 or resnet50-256-64 or darknet19"; do
  echo "=== $sel" >> $out/r2e_pytest.log
  python -m pytest tests -m gpu -v -s -k "$sel" -rs >> $out/r2e_pytest.log 2>&1; echo "pytest [$sel] rc=$?"
done
grep -E "PASSED|FAILED|ERROR|passed|failed|Abort|error" $out/r2e_pytest.log | tail -40
for cfg in "yolo9000 544 64" "yolo9000 544 16"; do
  Y2_HEAD_GAIN=13 python tools/throughput.py $cfg 20 --layers >> $out/r2e_throughput.txt 2>&1
done
grep "^{\|layer  23\|layer  24" $out/r2e_throughput.txt
