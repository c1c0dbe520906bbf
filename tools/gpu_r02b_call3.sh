mkdir -p gpurun_out
python tools/gpu_s1_bench.py > gpurun_out/c3_s1_bench.txt 2>&1; echo "s1 bench rc=$?"; cat gpurun_out/c3_s1_bench.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_s1_fused -s 5 -c 1 -f -o /tmp/c3_s1 python tools/gpu_s1_bench.py 384 > gpurun_out/c3_ncu_s1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c3_ncu_s1.log
python tools/ncu_summary.py /tmp/c3_s1.ncu-rep gpurun_out/c3_ncu_s1_fused.csv > /dev/null 2>&1; cat gpurun_out/c3_ncu_s1_fused.csv
ncu -i /tmp/c3_s1.ncu-rep --page details --csv 2>/dev/null | grep -E "Duration|Stall|L2 Hit|Throughput|Registers|Issue Slots|Eligible|Warp Cycles Per Issued|Bank" | cut -d, -f4-7 | cut -c1-200 | head -50 > gpurun_out/c3_ncu_s1_details.txt; cat gpurun_out/c3_ncu_s1_details.txt
ncu -i /tmp/c3_s1.ncu-rep --page source --csv > gpurun_out/c3_ncu_s1_source.csv 2>/dev/null; wc -l gpurun_out/c3_ncu_s1_source.csv
