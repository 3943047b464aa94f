#!/bin/bash
# ncu evidence of the final state: launch list of the headline step, full capture of the 256-channel pair tiles.  Outputs: gpurun_out/r7b_*
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r7b_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r7b_launches.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_conv3x3_pair<256" -s 2 -c 8 -f -o $O/r7b_ncu_pair256 \
    python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r7b_ncu_pair256.err
ls -la $O/r7b_* | awk '{print $5, $9}'
