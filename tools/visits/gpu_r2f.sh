#!/bin/bash
# full GPU tests, then PDL A/B in both regimes, then the ncu full capture of the dominant kernel (L23)
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -q -x -rs > $out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2f_pytest.log
for r in 1 2; do
  for mode in nopdl pdl; do
    if [ $mode = nopdl ]; then export Y2_NO_PDL=1; else unset Y2_NO_PDL; fi
    python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode sustained', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_u8']['value'], d['roofline']['frac'])"
    python bench.py --no-cpu-baseline --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode burst', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_u8']['value'], d['roofline']['frac'])"
  done
done
unset Y2_NO_PDL
ncu --set full --clock-control none --import-source on -k regex:conv_pair -c 2 -o $out/r2f_pair_L23 python tools/conv_bench.py --only L23 --reps 2 --warmup 1 > $out/r2f_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la $out/r2f_pair_L23.ncu-rep
