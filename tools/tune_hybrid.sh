#!/bin/bash
# developer tool: DRAM traffic of the X-store / means-load cache-hint variants
for m in 0 1 2 4 5 6; do
  lib=tools/lib_sm$m.so; [ $m = 0 ] && lib=prosstt_b200/libprosstt_b200.so
  echo "== store mode $m"
  PST_LIB=$lib ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:draw_counts_hybrid -s 1 -c 1 python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 1 2>&1 | grep -E "dram__|gpu__time|hit_rate"
  PST_LIB=$lib python tools/sampler_bench.py --cells 200000 --samplers hybrid --reps 3 2>&1 | grep -E "^hybrid"
done
