#!/bin/bash
# Evidence for the CTA-pair convolution kernel and the conv1_1 changes (one B200).  Outputs: gpurun_out/r7_*
O=gpurun_out
timeout 200 python -m pytest tests/test_conv3x3_gpu.py -x -q -k "cta_pair" 2>&1 | tail -2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_conv3x3_pair|k_conv1_gemm|k_conv1_stencil" -s 3 -c 30 -f -o $O/r7_ncu_full \
    python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r7_ncu_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r7_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r7_launches.err
ls -la $O/r7_* | awk '{print $5, $9}'
