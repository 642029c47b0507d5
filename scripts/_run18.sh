python scripts/ncu_cin.py 8192 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cin_tc -s 14 -c 7 -o gpurun_out/prof_r2_cin -f python scripts/ncu_cin.py 8192 > gpurun_out/ncu_cin.log 2>&1
tail -3 gpurun_out/ncu_cin.log
