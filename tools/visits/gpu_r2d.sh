#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q -k "multi_gpu or validation_files or reference_detector_golden or do_nms or sparse_tree or resnet50-256-64 or yolo9000" -rs > $out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/r2d_pytest.log
for cfg in "yolo9000 544 64" "yolo9000 544 16" "resnet50 256 64"; do
  Y2_HEAD_GAIN=13 python tools/throughput.py $cfg 20 --layers >> $out/r2d_throughput.txt 2>&1
done
grep "^{" $out/r2d_throughput.txt
