for d in 0 1 2 3 16 32 8; do
  echo "== dbg $d"; SGZ_T2_DBG=$d SGZ_CORR_TC_PROF=1 T2_REPS=3 T2_SIZES=6000 python tools/t2_size_probe.py 2>&1 | grep -v "^$" | tail -3 | cut -c1-400
done
echo "== narrow 0"; SGZ_T2_NARROW=0 SGZ_CORR_TC_PROF=1 T2_REPS=3 T2_SIZES=6000 python tools/t2_size_probe.py 2>&1 | grep -v "^$" | tail -3 | cut -c1-400
nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm,clocks.sm,power.draw --format=csv
