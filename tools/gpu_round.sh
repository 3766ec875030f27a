#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench lines (sustained default, the driver's 20-step command, the
# reference arm) and the ncu launch list of the bench command.  Usage: tools/gpu_round.sh <tag> [pytest args]
set -u
tag=${1:-r2a}; shift || true
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q -rs --durations=15 -s "$@" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/${tag}_smoke.log
python bench.py --profile-out $out/${tag}_layer_times.json > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_burst.json 2>> $out/${tag}_bench.err; echo "bench burst rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err; echo "ref arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --no-cpu-baseline --steps 4 --warmup 3 > $out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py $out/${tag}_launches.csv > $out/${tag}_launches_summary.txt 2>&1
head -c 1500 $out/${tag}_bench.json; echo
