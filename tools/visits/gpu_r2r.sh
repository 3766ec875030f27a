#!/bin/bash
# round 2 visit r: what bounds the slab kernel on 1x1 / narrow layers: full (0), no epilogue work (2), no MMAs (4),
# neither (6), everything but the TMA store issue (8)
out=gpurun_out; mkdir -p $out
for d in ${DBGS:-0 2 4 6 8}; do
  echo "== Y2_SLAB_DBG=$d"; Y2_SLAB_DBG=$d Y2_CONV_VARIANT=slab timeout 300 python tools/conv_bench.py --only L4,L5,L9,L13,L19,L26 --reps 40
done 2>&1 | tee $out/r2r_slab_decompose_${TAG:-a}.txt
