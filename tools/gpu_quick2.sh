mkdir -p gpurun_out/q
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/q/bench.json 2> gpurun_out/q/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/q/bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
