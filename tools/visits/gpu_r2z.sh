#!/bin/bash
# round 2 visit z: which 1x1 layers belong on the slab kernel now that its epilogue leaves through TMA stores
out=gpurun_out; mkdir -p $out
for v in "Y2_X=1" "Y2_SLAB_1X1_MIN_TILES=148" "Y2_SLAB_1X1_MIN_TILES=74" "Y2_X=2" "Y2_SLAB_1X1_MIN_TILES=148"; do
  echo "== $v"
  env $v timeout 300 python tools/throughput.py yolo-voc 416 64 20 | head -1 | cut -c1-200
  env $v timeout 300 python tools/throughput.py yolo-voc 416 64 400 | head -1 | cut -c1-120
  env $v timeout 300 python tools/throughput.py resnet50 256 64 20 | head -1 | cut -c1-200
  env $v timeout 300 python tools/throughput.py darknet19_448 448 64 20 | head -1 | cut -c1-120
done 2>&1 | tee $out/r2z_slab_1x1.txt
