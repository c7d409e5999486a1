nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown --format=csv,noheader -lms 100 > gpurun_out/clk.csv &
SMI=$!
timeout 100 python profiles/microbench.py --hq 32 --hk 32 --nkv 2048 --nq 2048 --causal --qf16 --steps 40000
timeout 100 python profiles/microbench.py --hq 32 --hk 32 --nkv 8192 --nq 8192 --causal --qf16 --steps 3000
kill $SMI
sort gpurun_out/clk.csv | uniq -c | sort -rn | head -12
