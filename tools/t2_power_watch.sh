# power draw / clocks / throttle reasons sampled every 100 ms while k_corr_tc2 scans the 1000 h database back to back
(T2_REPS=600 T2_SIZES=6000 python tools/t2_size_probe.py > gpurun_out/power_watch_scan.log 2>&1) &
P=$!
sleep 1
timeout 60 nvidia-smi --query-gpu=timestamp,power.draw,power.draw.instant,clocks.sm,clocks.mem,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,utilization.gpu --format=csv -lms 100 > gpurun_out/power_watch.csv &
S=$!
wait $P
kill $S
cat gpurun_out/power_watch_scan.log
