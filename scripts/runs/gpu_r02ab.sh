#!/bin/bash
mkdir -p gpurun_out/r02ab
timeout 300 python scripts/micro/pcie_duplex.py 4 > gpurun_out/r02ab/pcie.json 2> gpurun_out/r02ab/pcie.err
cat gpurun_out/r02ab/pcie.json
nvidia-smi topo -m > gpurun_out/r02ab/topo.txt 2>&1
numactl -H > gpurun_out/r02ab/numa.txt 2>&1 || lscpu | grep -i numa > gpurun_out/r02ab/numa.txt
