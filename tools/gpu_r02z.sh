#!/bin/bash
# round 2, GPU pass z: randomised cross-check of all device paths (tools/gpu_stress.py)
O=gpurun_out/r02z; mkdir -p $O
timeout 600 python tools/gpu_stress.py 1000 240 > $O/stress.log 2>&1; echo "stress rc=$?"; tail -3 $O/stress.log
