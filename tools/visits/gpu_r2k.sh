#!/bin/bash
# round 2 visit k: pair kernel with balanced variable-width pieces: kernel parity tests, A/B per layer, cost table
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "conv" > $out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2k_pytest.log
L=L8,L12,L18,L23,L29
{
for v in "Y2_PAIR_BALANCE=0" "Y2_PAIR_BALANCE=1" "Y2_PAIR_BALANCE=0 Y2_PAIR_FORCE_UNITS=1" "Y2_PAIR_BALANCE=0 Y2_PAIR_FORCE_UNITS=2" "Y2_PAIR_BALANCE=0 Y2_PAIR_FORCE_UNITS=3"; do
  echo "== $v"; env $v timeout 300 python tools/conv_bench.py --only $L --reps 40
done
} 2>&1 | tee $out/r2k_pair_balance.txt
for v in "Y2_PAIR_BALANCE=0" "Y2_PAIR_BALANCE=1"; do
  echo "== step $v"; env $v timeout 300 python tools/throughput.py yolo-voc 416 64 20 | head -1
  env $v timeout 300 python tools/throughput.py yolo-voc 416 64 400 | head -1
done 2>&1 | tee $out/r2k_step.txt
