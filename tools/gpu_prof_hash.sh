python tools/profile_hash.py > gpurun_out/r1_hash_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:itq_hash_tc -s 3 -c 1 -o gpurun_out/r1_hash_tc_full python tools/profile_hash.py > gpurun_out/r1_hash_ncu.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r1_hash_plain.log
