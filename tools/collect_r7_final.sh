#!/bin/bash
# Final evidence of this session (one B200): GPU test suite, headline bench line with its extras, reference arm, the other
# BASELINE configurations, full-model lines.  Outputs: gpurun_out/r7_*
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r7_gpu_tests.log 2>&1; tail -3 $O/r7_gpu_tests.log
timeout 600 python bench.py > $O/r7_bench_n1.json 2> $O/r7_bench_n1.err
timeout 300 python bench.py --impl reference > $O/r7_bench_reference_arm.json 2>/dev/null
for f in 13776 55104; do timeout 200 python bench.py --faces $f --no-extras --no-cpu-baseline > $O/r7_bench_cfg_$f.json 2>/dev/null; done
timeout 200 python bench.py --faces 220416 --img 540 --no-extras --no-cpu-baseline > $O/r7_bench_cfg_220416.json 2>/dev/null
ls -la $O/r7_bench* | awk '{print $5, $9}'
