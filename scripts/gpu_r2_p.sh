#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; head -c 300 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 300 python scripts/timeline_graph.py > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
