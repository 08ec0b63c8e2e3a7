mkdir -p gpurun_out
( for nw in 1 2 4; do timeout 60 tools/ubench/umma_bench 1 1 $nw; done ; echo "--- tmem A, 2 warps"; timeout 60 tools/ubench/umma_bench 1 2 2 ) > gpurun_out/umma_bench2.log 2>&1
cat gpurun_out/umma_bench2.log
