#!/bin/bash
# round 2 visit ae: classifier tail (avgpool with loads in flight, block-per-row softmax)
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_classifier_gpu.py -q -x -k "softmax or avgpool or classifier" > $out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2ae_pytest.log
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_golden_gpu.py -q -x -k "resnet or darknet19 or golden or alexnet" > $out/r2ae_pytest_net.log 2>&1; echo "pytest net rc=$?"; tail -3 $out/r2ae_pytest_net.log
{
python tools/throughput.py resnet50 256 64 20 --layers 2>&1 | grep -E "images_per_s|layer  6[6789]" | cut -c1-140
python tools/throughput.py darknet19_448 448 64 20 --layers 2>&1 | grep -E "images_per_s|layer  2[3456]" | cut -c1-140
} | tee $out/r2ae_classifier_tail.txt
