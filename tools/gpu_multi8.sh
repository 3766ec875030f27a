#!/bin/bash
# 8-GPU visit (charged 8x): the torchrun bench line at 8 ranks, the C multi-GPU entry, BASELINE config 3
out=gpurun_out; mkdir -p $out; tag=${1:-r2m8}
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $out/${tag}_bench_8gpu_burst.json 2> $out/${tag}_bench_8gpu.err; echo "bench burst rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 300 --warmup 30 > $out/${tag}_bench_8gpu.json 2>> $out/${tag}_bench_8gpu.err; echo "bench rc=$?"
head -c 400 $out/${tag}_bench_8gpu.json; echo
python tools/multi_bench.py --gpus 8 --steps 200 > $out/${tag}_multi_bench_8gpu.json 2> $out/${tag}_multi_bench.err; echo "multi_bench rc=$?"; cat $out/${tag}_multi_bench_8gpu.json
python tools/multi_bench.py --gpus 8 --cfg yolo --side 608 --batch 32 --head-gain 24 --steps 100 > $out/${tag}_multi_bench_c3_8gpu.json 2>> $out/${tag}_multi_bench.err; cat $out/${tag}_multi_bench_c3_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $out/${tag}_bench_8gpu_reference_arm.json 2>> $out/${tag}_bench_8gpu.err; echo "ref arm rc=$?"; cat $out/${tag}_bench_8gpu_reference_arm.json | cut -c1-300
