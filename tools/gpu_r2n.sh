#!/bin/bash
# round 2 visits n and ad (final refresh): ncu evidence - per-kernel metrics over whole steps (tensor-pipe % for the convolutions, DRAM bytes and
# duration for the bandwidth kernels) for yolo-voc b64 and yolo9000 b64, and the full capture of the dominant kernel
out=gpurun_out; mkdir -p $out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum,sm__inst_executed.sum,sm__cycles_active.avg"
ncu --metrics $M --clock-control none --launch-skip 160 -c 70 --csv --log-file $out/r2ad_step_metrics_yolo_voc.csv \
    python tools/throughput.py yolo-voc 416 64 2 > $out/r2ad_ncu_step.log 2>&1; echo "ncu step rc=$?"
Y2_HEAD_GAIN=13 ncu --metrics $M --clock-control none --launch-skip 130 -c 60 --csv --log-file $out/r2ad_step_metrics_yolo9000.csv \
    python tools/throughput.py yolo9000 544 64 2 > $out/r2ad_ncu_y9k.log 2>&1; echo "ncu y9k rc=$?"
ncu --metrics $M --clock-control none --launch-skip 300 -c 75 --csv --log-file $out/r2ad_step_metrics_resnet50.csv \
    python tools/throughput.py resnet50 256 64 2 > $out/r2ad_ncu_r50.log 2>&1; echo "ncu r50 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_pair -c 2 -o $out/r2ad_pair_L23 python tools/conv_bench.py --only L23 --reps 2 --warmup 1 > $out/r2ad_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/r2ad_pair_L23.ncu-rep --page raw --csv > $out/r2ad_pair_L23_raw.csv 2>/dev/null
ls -la $out/
