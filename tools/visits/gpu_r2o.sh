#!/bin/bash
# round 2 visit o: resident-weights one-tap pair kernel (1x1 layers): parity tests, per-layer A/B, step A/B
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv" > $out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2o_pytest.log
{
for v in "Y2_PAIR_NO_RESIDENT=1" "Y2_X=1"; do
  echo "== $v"; env $v timeout 300 python tools/conv_bench.py --only L9,L13,L19,L26 --reps 40
done
for v in "Y2_PAIR_NO_RESIDENT=1" "Y2_X=1"; do
  echo "== step $v"; env $v timeout 300 python tools/throughput.py yolo-voc 416 64 20 | head -1
  env $v timeout 300 python tools/throughput.py yolo-voc 416 64 400 | head -1
  env $v timeout 300 python tools/throughput.py resnet50 256 64 20 | head -1
  env $v timeout 300 python tools/throughput.py darknet19_448 448 64 20 | head -1
done
} 2>&1 | tee $out/r2o_resident.txt
timeout 600 python -m pytest tests/test_network_gpu.py -q -x -k "layer_activations" > $out/r2o_pytest_net.log 2>&1; echo "pytest net rc=$?"; tail -3 $out/r2o_pytest_net.log
