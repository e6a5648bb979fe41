timeout 100 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/dyn_tests.log
for i in 1 2; do
timeout 60 python bench.py --headline-only --no-cpu-baseline > gpurun_out/dyn_on_$i.json 2> gpurun_out/dyn_on_$i.err
timeout 60 python bench.py --headline-only --no-cpu-baseline --no-dyn-stream > gpurun_out/dyn_off_$i.json 2> gpurun_out/dyn_off_$i.err
done
python -c "
import json
for n in ('on_1','off_1','on_2','off_2'):
    d=json.load(open('gpurun_out/dyn_%s.json'%n)); print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('train_step_with_optimizer',{}).get('ms_per_step'))
"
timeout 60 python scripts/timeline_graph.py > gpurun_out/timeline_dyn.log 2>&1; cp gpurun_out/timeline.txt gpurun_out/timeline_dyn.txt
