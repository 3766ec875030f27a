#!/bin/bash
# developer helper: run conv_bench for several env configurations back to back on one box
# usage: tools/sweep.sh "L2,L4,L18" "VAR=a VAR2=b" "VAR=c" ...
layers=$1; shift
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader
for rep in 1 2; do
  for cfg in "$@"; do
    echo "== [$rep] $cfg"
    env $cfg python tools/conv_bench.py --only $layers 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('   %-4s %8.4f ms %7.1f TF' % (d['layer'], d['ms'], d['tflops']))
"
  done
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader
