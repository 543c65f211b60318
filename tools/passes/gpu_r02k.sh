#!/bin/bash
# round 2, GPU pass k: TMA-fed header ring in the batch bit kernel (on/off), initcheck after the fixes
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "bit or latency or config3 or config2" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
P="python tools/gpu_probe.py --no-imad --reps 5"
for R in 0 1; do for B in 16384 65536; do echo "== ring $R" >> $O/probe_sha.jsonl; GW_BIT_RING=$R timeout 300 $P --circuits circuit8_sha256_512 --batch $B >> $O/probe_sha.jsonl 2>> $O/probe.err; done; done
cut -c1-110 $O/probe_sha.jsonl
for R in 0 1; do echo "== ring $R" >> $O/probe_small.jsonl; GW_BIT_RING=$R timeout 300 $P --circuits circuit6_num2bits --batch 65536 >> $O/probe_small.jsonl 2>> $O/probe.err; GW_BIT_RING=$R timeout 300 $P --circuits circuit6_num2bits --batch 1048576 >> $O/probe_small.jsonl 2>> $O/probe.err; done
cut -c1-110 $O/probe_small.jsonl
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?" | tee -a $O/sanitizer_initcheck.log; grep -c Uninit $O/sanitizer_initcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_probe.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/sanitizer_racecheck.log; grep -c "Race reported" $O/sanitizer_racecheck.log
