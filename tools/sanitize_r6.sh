#!/bin/bash
# compute-sanitizer over the kernels written in the second half of round 2 (mesh tile kernels, weight-gradient GEMM, mesh_prep,
# posenc, rodrigues, narrow linear, shadow background row, conv1_1 mask, block-aggregated raster binning).
# Usage (on a GPU box): bash tools/sanitize_r6.sh
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SEL_MEM="tests/test_mesh_gpu.py tests/test_model_gpu.py::test_rodrigues_kernel_matches_torch_float64 tests/test_raster_gpu.py tests/test_shadow_gpu.py"
timeout 1500 compute-sanitizer --tool memcheck python -m pytest $SEL_MEM -q -x -p no:cacheprovider -k "not 120000 and not 30000 and not cuda_graph" > gpurun_out/r6_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/r6_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_modules_gpu.py tests/test_losses_gpu.py -q -x -p no:cacheprovider -k "4099 or 4096-128 or 5008 or 48-32 or 16016 or 4096-128-3 or 5001-64-1 or 4100-132-2 or (fused_lpips_matches_oracle and tcgen05 and hw0) or first_convolution_kernels" > gpurun_out/r6_sanitizer_memcheck_mlp.log 2>&1
tail -4 gpurun_out/r6_sanitizer_memcheck_mlp.log
timeout 1200 compute-sanitizer --tool racecheck python -m pytest "tests/test_mesh_gpu.py::test_rasterize_mesh_forward_backward_matches_oracle" tests/test_modules_gpu.py tests/test_raster_gpu.py tests/test_shadow_gpu.py -q -x -p no:cacheprovider -k "2000-size0 or 2000-size3 or 48-32 or 4096-128 or 5008 or interleaved or ragged or deterministic" > gpurun_out/r6_sanitizer_racecheck.log 2>&1
tail -4 gpurun_out/r6_sanitizer_racecheck.log
