#!/bin/bash
# round 2 visit v: one-tap pair kernel for wide fp32 heads (yolo9000)
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "f32_flat" > $out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2v_pytest.log
for v in "Y2_PAIR_NO_F32_HEAD=1" "Y2_X=1"; do
  echo "== $v"
  env $v Y2_HEAD_GAIN=13 timeout 300 python tools/throughput.py yolo9000 544 64 20 --layers 2>&1 | grep -E "images_per_s|layer  23" | cut -c1-140
  env $v Y2_HEAD_GAIN=13 timeout 300 python tools/throughput.py yolo9000 544 128 10 | head -1 | cut -c1-140
done 2>&1 | tee $out/r2v_y9k_head.txt
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_golden_gpu.py tests/test_baseline_batches_gpu.py -q -x -k "yolo9000 or tree" > $out/r2v_pytest_net.log 2>&1; echo "pytest net rc=$?"; tail -3 $out/r2v_pytest_net.log
