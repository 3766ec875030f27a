#!/bin/bash
# round 2 visit y: staged fp32 stores for the narrow detection heads (125 / 425 filters) on the slab kernel
out=gpurun_out; mkdir -p $out
for v in "Y2_SLAB_NO_F32_STAGE=1" "Y2_X=1"; do
  echo "== $v"
  env $v timeout 300 python tools/throughput.py yolo-voc 416 64 20 --layers 2>&1 | grep -E "images_per_s|layer  30" | cut -c1-140
  env $v timeout 300 python tools/throughput.py yolo 608 32 20 --layers 2>&1 | grep -E "images_per_s|layer  30" | cut -c1-140
  env $v timeout 300 python tools/throughput.py tiny-yolo-voc 416 64 20 --layers 2>&1 | grep -E "images_per_s|layer  14" | cut -c1-140
done 2>&1 | tee $out/r2y_heads.txt
timeout 1200 python -m pytest tests/test_network_gpu.py tests/test_golden_gpu.py tests/test_kernels_gpu.py tests/test_detector_cpp.py tests/test_validate_gpu.py tests/test_demo_gpu.py -q -x > $out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2y_pytest.log
