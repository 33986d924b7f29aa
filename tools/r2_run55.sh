#!/bin/bash
# FINAL end state of round 2 (loop graph, group-local norm, in-place stride-2 conv, upsample fold): full GPU suite, smoke, bench line, ncu launch lists (+ DRAM traffic) of the timed plans
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2_final_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/r2_final_gputests.log
timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 200 gpurun_out/r2_final_bench.json
python tools/profile_unet.py unet short > gpurun_out/prof_plain_unet.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches_unet.csv python tools/profile_unet.py unet short > gpurun_out/prof_ncu_unet.log 2>&1
python tools/profile_unet.py vae > gpurun_out/prof_plain_vae.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches_vae.csv python tools/profile_unet.py vae > gpurun_out/prof_ncu_vae.log 2>&1
python tools/join_launches.py gpurun_out/r2_final_launches_unet.csv gpurun_out/launch_descs_unet.txt 70 --traffic-json gpurun_out/unet_gemm_traffic.json > gpurun_out/r2_final_launch_table_unet.txt
python tools/join_launches.py gpurun_out/r2_final_launches_vae.csv gpurun_out/launch_descs_vae.txt 40 > gpurun_out/r2_final_launch_table_vae.txt
head -10 gpurun_out/r2_final_launch_table_unet.txt; head -6 gpurun_out/r2_final_launch_table_vae.txt; cat gpurun_out/unet_gemm_traffic.json
