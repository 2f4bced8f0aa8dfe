#!/bin/bash
# Memory check of the kernels on the HOST: the tests that execute CUDA sources through tests/cuda_emu, rebuilt with
# AddressSanitizer and run with libasan preloaded, so an out-of-bounds load or store of a kernel into a heap buffer
# (torch / numpy allocations get redzones once malloc is intercepted) aborts the test.  The CPU-side counterpart of
# `compute-sanitizer --tool memcheck`; usage: bash tools/emu_memcheck.sh [pytest -k expression]
cd "$(dirname "$0")/.."
ASAN_LIB=$(gcc -print-file-name=libasan.so)
export GVD_EMU_FULL=1 GVD_EMU_ASAN=1 ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0:abort_on_error=1 LD_PRELOAD="$ASAN_LIB"
python -m pytest tests/test_nn_bwd_emu_cpu.py tests/test_nn_fwd_emu_cpu.py tests/test_pcd2img_cpu.py tests/test_train_ops_cpu.py \
    tests/test_knn_cpu.py tests/test_raster_emu_cpu.py tests/test_tattn_mma_emu_cpu.py tests/test_gemm_emu_cpu.py \
    tests/test_attn_fwd_emu_cpu.py tests/test_attn_bwd_emu_cpu.py tests/test_unet_emu_cpu.py -q -x -W ignore -p no:cacheprovider ${1:+-k "$1"}
