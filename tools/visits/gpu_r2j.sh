#!/bin/bash
# round 2 visit j: sensitivity of the pair kernel to weight-tile traffic (Y2_PAIR_DBG=1 skips the weight loads of odd
# taps - wrong results, timing only), all-config throughput record
out=gpurun_out; mkdir -p $out
for v in "" "Y2_PAIR_DBG=1"; do
  echo "== pair kernel $v"; env $v python tools/conv_bench.py --only L8,L12,L18,L23,L29 --reps 40
done 2>&1 | tee $out/r2j_pair_halfB.txt
{
python tools/throughput.py tiny-yolo-voc 416 64 20 | head -1
python tools/throughput.py yolo-voc 416 64 20 | head -1
python tools/throughput.py yolo 608 32 20 | head -1
python tools/throughput.py yolo 608 64 20 | head -1
Y2_HEAD_GAIN=13 python tools/throughput.py yolo9000 544 16 20 | head -1
Y2_HEAD_GAIN=13 python tools/throughput.py yolo9000 544 64 20 | head -1
Y2_HEAD_GAIN=13 python tools/throughput.py yolo9000 544 128 10 | head -1
python tools/throughput.py darknet19_448 448 64 20 | head -1
python tools/throughput.py darknet19_448 448 512 5 | head -1
python tools/throughput.py resnet50 256 64 20 | head -1
python tools/throughput.py resnet50 256 256 20 | head -1
python tools/throughput.py resnet50 256 512 10 | head -1
} > $out/r2j_throughput_all_configs.jsonl 2> $out/r2j_throughput.err
cat $out/r2j_throughput_all_configs.jsonl
python tools/throughput.py resnet50 256 64 20 --layers > $out/r2j_resnet50_layers.txt 2>&1
