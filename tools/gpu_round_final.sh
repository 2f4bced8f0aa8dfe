#!/bin/bash
# Final artefacts of round 2 (one GPU): both bench arms, ncu launch list + full raster capture + traffic, U-Net / guided kernel breakdowns.
./tools/gpu_profile_round.sh r02f
timeout 600 python tools/profile_guided.py unet > gpurun_out/r02f_unet_kernel_breakdown.txt 2>&1
timeout 600 python tools/profile_guided.py guided host > gpurun_out/r02f_guided_kernel_breakdown.txt 2>&1
timeout 300 python tools/profile_host.py > gpurun_out/r02f_raster_host_profile.txt 2>&1
head -3 gpurun_out/r02f_unet_kernel_breakdown.txt gpurun_out/r02f_guided_kernel_breakdown.txt | cut -c1-120
