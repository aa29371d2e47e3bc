#!/bin/sh
# which OpenMP thread count is best for the reference CPU V-cycle on this host?
nproc; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)" 
for t in 8 16 32 64 128; do
  echo "threads $t: $(OMP_NUM_THREADS=$t OMP_PROC_BIND=${BIND:-false} python oracle/cpu_baseline.py --level 4 --warmup 1 --steps 3 | cut -c1-90)"
done
