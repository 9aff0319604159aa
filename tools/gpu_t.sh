python tools/l2_prof.py > gpurun_out/l2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:itq_hash_tc_kernel -s 14 -c 1 -o gpurun_out/r1_l2_filter_full python tools/l2_prof.py > gpurun_out/l2_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/l2_ncu.log
