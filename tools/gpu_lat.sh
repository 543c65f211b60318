mkdir -p gpurun_out/lat
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "latency or poseidon_like" 2>&1 | tail -3
timeout 300 python tools/gpu_latency.py --reps 20 2>&1 | tee gpurun_out/lat/latency.jsonl
