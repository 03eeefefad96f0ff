#!/bin/bash
# DRAM traffic of the KV-projection GEMM vs N-group budget x L2 eviction hints (A loads, W loads, C stores)
for CFG in "70 fnf" "70 flf" "24 fnf" "36 fnf"; do
  set -- $CFG
  echo "=== TDC_GEMM_NGROUP_MB=$1 TDC_GEMM_HINTS=$2 (A,W,C)"
  TDC_GEMM_NGROUP_MB=$1 TDC_GEMM_HINTS=$2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:tdc_gemm -c 1 ./build/gemm_test 2 370800 9216 3584 0 1 2>&1 | grep -E "dram__bytes|gpu__time|hit_rate"
done
# query-side shapes: does evict_first on C hurt the next kernel that reads C?  (timing only, no ncu)
for H in nln fnf; do echo "=== hints $H"; TDC_GEMM_HINTS=$H ./build/gemm_test 2 86400 2304 768 0 20 | grep time; TDC_GEMM_HINTS=$H ./build/gemm_test 2 86400 768 768 2 20 | grep time; TDC_GEMM_HINTS=$H ./build/gemm_test 2 86400 3072 768 1 20 | grep time; done
