#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
bash tools/gpu_scale2.sh $tag 1 2 4 8
timeout 300 tools/microbench/pcie_multi 256 8 > gpurun_out/${tag}_pcie_multi.txt 2>&1; echo "pcie_multi rc=$?"; cat gpurun_out/${tag}_pcie_multi.txt
