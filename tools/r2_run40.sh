#!/bin/bash
# evidence of the round-2 end state: bench line, ncu launch lists (+ DRAM traffic) of the timed plans, ncu --set full of the new kernels
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -c 600 gpurun_out/r2s_bench.json
python tools/profile_unet.py unet short > gpurun_out/prof_plain_unet.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_unet.csv python tools/profile_unet.py unet short > gpurun_out/prof_ncu_unet.log 2>&1
tail -2 gpurun_out/prof_ncu_unet.log
python tools/profile_unet.py vae > gpurun_out/prof_plain_vae.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_vae.csv python tools/profile_unet.py vae > gpurun_out/prof_ncu_vae.log 2>&1
tail -2 gpurun_out/prof_ncu_vae.log
python tools/join_launches.py gpurun_out/r2s_launches_unet.csv gpurun_out/launch_descs_unet.txt 70 --traffic-json gpurun_out/unet_gemm_traffic.json > gpurun_out/r2s_launch_table_unet.txt
python tools/join_launches.py gpurun_out/r2s_launches_vae.csv gpurun_out/launch_descs_vae.txt 40 > gpurun_out/r2s_launch_table_vae.txt
head -14 gpurun_out/r2s_launch_table_unet.txt; head -8 gpurun_out/r2s_launch_table_vae.txt; cat gpurun_out/unet_gemm_traffic.json
python tools/gn_parts_one.py > /dev/null 2>&1 && \
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'gn_parts_kernel|gemm_tc_kernel' -f -o gpurun_out/r2s_gn_parts python tools/gn_parts_one.py > gpurun_out/r2s_gn_parts_ncu.log 2>&1
tail -3 gpurun_out/r2s_gn_parts_ncu.log; ls -la gpurun_out/*.ncu-rep
