#!/bin/bash
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q -x -k "kernels_gpu or reorg or resnet50 or serving or gemm or im2col or yolo-1-608 or yolo9000-1" > $out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2h_pytest.log
for v in "" "Y2_CONV_VARIANT=pertap"; do
  echo "== wide shapes $v"; env $v python tools/conv_bench.py --batch 32 --only W4_608,W4_544,W8_608
done
python tools/throughput.py resnet50 256 64 20 --layers > $out/r2h_resnet50.txt 2>&1; head -1 $out/r2h_resnet50.txt; grep "layer   0\|layer  15\|layer  31\|layer  55" $out/r2h_resnet50.txt
python tools/throughput.py resnet50 256 256 20 | head -1
python tools/throughput.py yolo 608 32 20 | head -1
python tools/throughput.py yolo 608 64 20 | head -1
Y2_HEAD_GAIN=13 python tools/throughput.py yolo9000 544 64 20 | head -1
Y2_HEAD_GAIN=13 python tools/throughput.py yolo9000 544 128 10 | head -1
python tools/throughput.py darknet19_448 448 64 20 | head -1
python tools/throughput.py tiny-yolo-voc 416 64 20 | head -1
