#!/bin/bash
# developer tool: compute-sanitizer over the count kernels and the lineage/index kernels (small sizes)
for tool in memcheck racecheck; do
  echo "== $tool: sampler_bench (both samplers, 1500 cells x 20000 genes)"
  compute-sanitizer --tool $tool --print-limit 5 python tools/sampler_bench.py --cells 1500 --reps 1 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|^hybrid|^gamma" | head -12
done
echo "== memcheck + racecheck: fused per-gene summaries (STATS instantiations), staged host transports, tail fix-up, gamma_poisson pipeline at three depths"
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_gene_stats or staged_host or padded_row_stride or gamma_poisson_pipeline" 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|passed|failed|Invalid" | head
done
echo "== memcheck: lineage (level-batched loop, checks kernel) + index-map tests"
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "walks or index_maps or pick_branch or pearson or nb_params or domain or lineage or parameterisation" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head
echo "== memcheck: epilogue kernels (stats, transforms, CSR, narrow formats, base expression)"
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -k "count_stats or transform_counts or csr_compaction or narrow_kernel or base_gene_exp" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head
