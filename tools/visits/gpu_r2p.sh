#!/bin/bash
# round 2 visit p: what bounds the 1x1 layers - full ncu captures of L13 under the three kernels; PDL on the small kernels
out=gpurun_out; mkdir -p $out
cap() { # tag, env..., kernel regex
  tag=$1; shift; rx=$1; shift
  env "$@" ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 --launch-skip 2 -o $out/r2p_$tag python tools/conv_bench.py --only L13 --reps 2 --warmup 1 > $out/r2p_$tag.log 2>&1; echo "ncu $tag rc=$?"
  ncu -i $out/r2p_$tag.ncu-rep --page raw --csv > $out/r2p_${tag}_raw.csv 2>/dev/null
  ncu -i $out/r2p_$tag.ncu-rep --page source --csv > $out/r2p_${tag}_source.csv 2>/dev/null
}
cap L13_slab conv_slab Y2_PAIR_NO_RESIDENT=1
cap L13_pair_resident conv_pair Y2_X=1
cap L13_pair_stream conv_pair Y2_PAIR_NO_RESIDENT=1 Y2_CONV_VARIANT=pair
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_golden_gpu.py -q -x > $out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2p_pytest.log
for v in "Y2_NO_PDL=1" "Y2_X=1"; do
  echo "== step $v"; env $v Y2_PAIR_NO_RESIDENT=1 timeout 300 python tools/throughput.py yolo-voc 416 64 20 | head -1
  env $v Y2_PAIR_NO_RESIDENT=1 timeout 300 python tools/throughput.py yolo-voc 416 64 400 | head -1
  env $v Y2_PAIR_NO_RESIDENT=1 timeout 300 python tools/throughput.py resnet50 256 64 20 | head -1
done 2>&1 | tee $out/r2p_pdl.txt
ls -la $out | head -30
