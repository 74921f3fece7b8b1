#!/bin/bash
# build the standalone LU test drivers: scratch/build_ll.sh "13 8 5" [extra nvcc flags]
cd /root/repo/scratch
for nt in $1; do
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DNTV=$nt $2 ll_test.cu -o ll_test_$nt -Xptxas -v 2>&1 | grep -E "error|registers|spill" | tr '\n' ' '; echo "NT=$nt";
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DWFO_LL_CLOCKS -DNTV=$nt $2 ll_test.cu -o ll_test_clk_$nt 2>&1 | grep error ) &
done
wait
