#!/bin/bash
# Developer tool (GPU box): A/B of a plan-time switch in the SUSTAINED regime (bench.py default: 300 timed steps),
# alternating in one box.  usage: tools/ab_sustained.sh VAR=VALUE
for r in off on off on; do
  if [ $r = on ]; then unset ${1%%=*}; else export "$1"; fi
  python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$r' == 'on' and 'default' or '$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
