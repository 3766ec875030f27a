#!/bin/bash
# N-GPU visit: the multi-GPU tests, the torchrun bench line at N ranks and the C multi-GPU entry's e2e.
# Usage: tools/gpu_multi.sh <tag> <N> [steps]
tag=$1; N=$2; steps=${3:-300}
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
python -m pytest tests -m gpu -q -x -k "multi_gpu" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $steps --warmup 30 > $out/${tag}_bench_${N}gpu.json 2> $out/${tag}_bench_${N}gpu.err; echo "bench rc=$?"
head -c 600 $out/${tag}_bench_${N}gpu.json; echo
python tools/multi_bench.py --gpus $N --steps 200 > $out/${tag}_multi_bench_${N}gpu.json 2> $out/${tag}_multi_bench.err; echo "multi_bench rc=$?"; cat $out/${tag}_multi_bench_${N}gpu.json
Y2_NO_NUMA_BIND=1 python tools/multi_bench.py --gpus $N --steps 200 > $out/${tag}_multi_bench_${N}gpu_nobind.json 2>> $out/${tag}_multi_bench.err; cat $out/${tag}_multi_bench_${N}gpu_nobind.json
# BASELINE config 3: yolo.cfg 608, global batch 256 (32 per GPU at 8 GPUs, 64 at 4, 128 at 2)
python tools/multi_bench.py --gpus $N --cfg yolo --side 608 --batch $((256 / N)) --head-gain 24 --steps 60 > $out/${tag}_multi_bench_c3_${N}gpu.json 2>> $out/${tag}_multi_bench.err; cat $out/${tag}_multi_bench_c3_${N}gpu.json
