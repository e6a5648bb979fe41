#!/bin/bash
# end-of-round check on one GPU, as the driver runs it: GPU suite, smoke, reference arm, bench; plus the launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; head -c 300 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 300 python scripts/timeline_graph.py > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 1 --warmup 3 --rollout 200 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
