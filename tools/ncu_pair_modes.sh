#!/bin/bash
# Developer tool (GPU box): ncu counters of the pair kernel on one layer shape with whole tiles round-robin
# (Y2_PAIR_BALANCE=0) and with the balanced variable-width pieces, to see where the elapsed cycles go.
L=${1:-L23}
M="gpu__time_duration.sum,sm__cycles_elapsed.max,sm__cycles_active.avg,sm__pipe_tensor_cycles_active.avg,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,sm__inst_executed_pipe_tensor.sum"
for mode in 0 1; do
  Y2_PAIR_BALANCE=$mode ncu --clock-control none --metrics $M -k regex:conv_pair --csv \
     --log-file gpurun_out/ncu_pair_${L}_bal${mode}.csv python tools/conv_bench.py --only $L --reps 3 --warmup 2 > gpurun_out/ncu_pair_${L}_bal${mode}.log 2>&1
done
