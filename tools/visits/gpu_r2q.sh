#!/bin/bash
# round 2 visit q: lighter cluster arrive in the pair kernel, PDL on the small kernels (A/B in one box)
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_golden_gpu.py -q -x > $out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2q_pytest.log
timeout 300 python tools/conv_bench.py --only L8,L12,L18,L23,L29 --reps 40 2>&1 | tee $out/r2q_pair.txt
for r in 1 2; do
for v in "Y2_NO_PDL_SMALL=1" "Y2_X=1"; do
  echo "== step $v"; env $v timeout 300 python tools/throughput.py yolo-voc 416 64 20 | head -1 | cut -c1-120
  env $v timeout 300 python tools/throughput.py yolo-voc 416 64 400 | head -1 | cut -c1-120
  env $v timeout 300 python tools/throughput.py resnet50 256 64 20 | head -1 | cut -c1-120
  env $v timeout 300 python tools/throughput.py tiny-yolo-voc 416 64 20 | head -1 | cut -c1-120
done
done 2>&1 | tee $out/r2q_pdl_small.txt
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_demo_gpu.py tests/test_classifier_gpu.py -q -x > $out/r2q_pytest_net.log 2>&1; echo "pytest net rc=$?"; tail -3 $out/r2q_pytest_net.log
