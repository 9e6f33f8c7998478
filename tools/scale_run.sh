#!/bin/bash
# Multi-GPU measurements for N ranks of one box: concurrent PCIe probe, bench.py weak / strong scaling, C5 pulls.
# usage: tools/scale_run.sh N TAG      (outputs under gpurun_out/)
N=$1; TAG=$2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
if [ "$N" = "1" ]; then RUN="python"; fi
$RUN tools/pcie_probe.py > gpurun_out/${TAG}_pcie_$N.json 2> gpurun_out/${TAG}_pcie_$N.err
$RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_weak_$N.json 2> gpurun_out/${TAG}_weak_$N.err
$RUN bench.py --gpus $N --steps 10 --warmup 3 --scaling strong > gpurun_out/${TAG}_strong_$N.json 2> gpurun_out/${TAG}_strong_$N.err
$RUN bench.py --gpus $N --steps 5 --warmup 3 --workload c5 > gpurun_out/${TAG}_c5_$N.json 2> gpurun_out/${TAG}_c5_$N.err
for f in weak strong c5; do python tools/show_bench.py gpurun_out/${TAG}_${f}_$N.json || tail -3 gpurun_out/${TAG}_${f}_$N.err; done
cat gpurun_out/${TAG}_pcie_$N.json
