#!/bin/bash
# Evidence of the final state of round 2 (one B200): tests, bench lines, ncu launch lists and full captures.  Outputs: gpurun_out/r6_*
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r6_gpu_tests.log 2>&1; tail -3 $O/r6_gpu_tests.log
python bench.py > $O/r6_bench_n1.json 2> $O/r6_bench_n1.err
python bench.py --impl reference > $O/r6_bench_reference_arm.json 2>/dev/null
for f in 13776 55104; do python bench.py --faces $f --no-extras --no-cpu-baseline > $O/r6_bench_cfg_$f.json 2>/dev/null; done
python bench.py --faces 220416 --img 540 --no-extras --no-cpu-baseline > $O/r6_bench_cfg_220416.json 2>/dev/null
python bench.py --full-model --no-cpu-baseline > $O/r6_bench_full_model.json 2>/dev/null
python bench.py --full-model --frames-per-step 1 --no-cpu-baseline > $O/r6_bench_full_model_b1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r6_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r6_launches.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1400 --csv --log-file $O/r6_launches_full.csv \
    python bench.py --full-model --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r6_launches_full.err
ncu --set full --clock-control none --import-source on -k regex:"k_mesh_tiles_fwd|k_mesh_tiles_bwd|k_linear_wgrad|k_preprocess|k_emit|k_nonrigid_input_fwd" -s 12 -c 8 -f -o $O/r6_ncu_full \
    python bench.py --full-model --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r6_ncu_full.err
ls -la $O/r6_* | awk '{print $5, $9}'
