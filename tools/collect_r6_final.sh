#!/bin/bash
# Final evidence of round 2 (one B200): GPU test suite, headline bench line with its extras, full-model lines, ncu capture of the
# weight-gradient GEMM and the mesh backward / shadow background kernels.  Outputs: gpurun_out/r6_*
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r6_gpu_tests.log 2>&1; tail -3 $O/r6_gpu_tests.log
python bench.py > $O/r6_bench_n1.json 2> $O/r6_bench_n1.err
python bench.py --full-model --no-cpu-baseline > $O/r6_bench_full_model.json 2>/dev/null
python bench.py --full-model --frames-per-step 1 --no-cpu-baseline > $O/r6_bench_full_model_b1.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"k_linear_wgrad|k_mesh_tiles_bwd|k_shadow_bg|k_narrow_linear|k_nonrigid_input_bwd" -s 10 -c 14 -f -o $O/r6_ncu_full2 \
    python bench.py --full-model --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-cuda-graph > /dev/null 2> $O/r6_ncu_full2.err
ls -la $O/r6_ncu_full2.ncu-rep | awk '{print $5, $9}'
