#!/bin/bash
# round 2 visit s: division-free shortcut kernel: kernel test, resnet50 parity, throughput A/B
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "shortcut" > $out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2s_pytest.log
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_golden_gpu.py -q -x -k "resnet or golden" > $out/r2s_pytest_net.log 2>&1; echo "pytest net rc=$?"; tail -3 $out/r2s_pytest_net.log
for v in "Y2_SHORTCUT_GENERAL=1" "Y2_X=1"; do
  echo "== $v"; env $v timeout 300 python tools/throughput.py resnet50 256 64 20 | head -1 | cut -c1-130
  env $v timeout 300 python tools/throughput.py resnet50 256 256 20 | head -1 | cut -c1-130
done 2>&1 | tee $out/r2s_shortcut.txt
python tools/throughput.py resnet50 256 64 20 --layers 2>&1 | grep "type 13" | head -20 | tee -a $out/r2s_shortcut.txt
