#!/bin/bash
# round 2 visit x: staged fp32 heads from 512 filters (the 1000-way classifier heads) on the pair kernel
out=gpurun_out; mkdir -p $out
for v in "Y2_PAIR_NO_F32_HEAD=1 Y2_SLAB_NO_F32_STAGE=1" "Y2_X=1"; do
  echo "== $v"
  env $v timeout 300 python tools/throughput.py resnet50 256 64 20 --layers 2>&1 | grep -E "images_per_s|layer  66" | cut -c1-140
  env $v timeout 300 python tools/throughput.py darknet19_448 448 64 20 --layers 2>&1 | grep -E "images_per_s|layer  23" | cut -c1-140
done 2>&1 | tee $out/r2x_heads.txt
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_classifier_gpu.py tests/test_golden_gpu.py -q -x -k "resnet or darknet19 or classifier or golden or alexnet" > $out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2x_pytest.log
