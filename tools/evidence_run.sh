set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launches.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_corr_tc2 -s 3 -c 1 -f -o gpurun_out/r02_k_corr_tc2 python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_k1.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none -k regex:k_replay_fill_po -s 1 -c 1 -f -o gpurun_out/r02_k_replay_fill_po python tools/punchout_probe.py > gpurun_out/ncu_po.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none -k regex:k_segm_pick_warp -s 1 -c 1 -f -o gpurun_out/r02_k_segm_pick_warp python tools/segm_probe.py > gpurun_out/ncu_segm.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none -k regex:k_cross -s 1 -c 1 -f -o gpurun_out/r02_k_cross python tools/cross_probe.py > gpurun_out/ncu_cross.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none -k regex:k_stats_hist -s 1 -c 1 -f -o gpurun_out/r02_k_stats_hist python tools/stats_probe.py > gpurun_out/ncu_stats.log 2>&1; echo rc=$?
ls -la gpurun_out/
# k_corr_tc2 under sustained load: role counters, per-CTA wall clock and mean SM clock per ablation; power / clock trace
bash tools/t2_power_ablate.sh > gpurun_out/r02_k_corr_tc2_power_ablation.txt 2>&1; echo rc=$?
bash tools/t2_power_watch.sh; cp gpurun_out/power_watch.csv gpurun_out/r02_k_corr_tc2_power_watch.csv
python tools/t2_size_probe.py > gpurun_out/size_probe.log 2>&1; echo rc=$?
