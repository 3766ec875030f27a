#!/bin/bash
# round 2 visit t: detection lists on their own D2H stream, device-resident pipelined value loop
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_demo_gpu.py tests/test_detector_cpp.py -q -x -k "pipelined or u8 or frames or demo or detector or multi" > $out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2t_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2t_bench_burst.json 2> $out/r2t_bench.err; echo "bench burst rc=$?"
python bench.py --no-cpu-baseline > $out/r2t_bench.json 2>> $out/r2t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2t_bench_burst.json", "gpurun_out/r2t_bench.json"):
    d = json.load(open(f))
    print(f, "value", d["value"], d["ms_per_step"], "sync", d["config"]["sync_value"], "e2e", d["e2e"]["value"], "u8", d["e2e_u8"]["value"], "frac", d["roofline"]["frac"], d["model_frac_of_peak"])
PY
tail -5 $out/r2t_bench.err
