#!/bin/bash
# device-resident throughput of every BASELINE config (tools/throughput.py), one line per (cfg, batch)
tag=${1:-r2ab}; out=gpurun_out; mkdir -p $out
{
python tools/throughput.py tiny-yolo-voc 416 1 50 | head -1
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
} > $out/${tag}_throughput_all_configs.jsonl 2> $out/${tag}_throughput.err
cut -c1-150 $out/${tag}_throughput_all_configs.jsonl
