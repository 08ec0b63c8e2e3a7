mkdir -p gpurun_out
( for nw in 3 6 8; do for sd in 0 1; do timeout 60 tools/ubench/umma_bench 1 1 $nw $sd | grep -v "shift=1"; done; done ) > gpurun_out/umma_bench3.log 2>&1
cat gpurun_out/umma_bench3.log
