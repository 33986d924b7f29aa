#!/bin/bash
# ncu launch lists (time + DRAM bytes per launch) of the timed UNet plan and of the VAE decoder plan
mkdir -p gpurun_out
python tools/profile_unet.py unet short > gpurun_out/prof_plain_unet.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_unet.csv python tools/profile_unet.py unet short > gpurun_out/prof_ncu_unet.log 2>&1
tail -2 gpurun_out/prof_ncu_unet.log
python tools/profile_unet.py vae > gpurun_out/prof_plain_vae.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_vae.csv python tools/profile_unet.py vae > gpurun_out/prof_ncu_vae.log 2>&1
tail -2 gpurun_out/prof_ncu_vae.log
python tools/join_launches.py gpurun_out/r2_launches_unet.csv gpurun_out/launch_descs_unet.txt 70 --traffic-json gpurun_out/unet_gemm_traffic.json > gpurun_out/r2_launch_table_unet.txt
python tools/join_launches.py gpurun_out/r2_launches_vae.csv gpurun_out/launch_descs_vae.txt 40 > gpurun_out/r2_launch_table_vae.txt
head -12 gpurun_out/r2_launch_table_unet.txt; head -8 gpurun_out/r2_launch_table_vae.txt; cat gpurun_out/unet_gemm_traffic.json
