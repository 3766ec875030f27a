#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in "Y2_PAIR_DBG=16" "Y2_X=1"; do
  echo "== $v"
  env $v Y2_HEAD_GAIN=13 timeout 300 python tools/throughput.py yolo9000 544 64 20 --layers 2>&1 | grep -E "images_per_s|layer  23" | cut -c1-140
done 2>&1 | tee $out/r2w_y9k_stcs.txt
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum"
Y2_HEAD_GAIN=13 ncu --metrics $M --clock-control none -k regex:conv_pair_kernel.*1 -c 2 --launch-skip 3 --csv --log-file $out/r2w_y9k_head_ncu.csv python tools/throughput.py yolo9000 544 64 2 > $out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" $out/r2w_y9k_head_ncu.csv | cut -d, -f5,13- | tail -14
