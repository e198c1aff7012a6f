#!/bin/bash
# one kernel-iteration round on the GPU box: marching parity tests, variant timings, one ncu --set full capture
# usage: tools/gpu_iter.sh <tag> [gens] [ncu-gen]
tag=${1:-it}; gens=${2:-v1,v2,v2eo}; ncugen=${3:-v2eo}
mkdir -p gpurun_out
python -m pytest tests/test_pressure_gpu.py -x -q -k "marching" 2>&1 | tail -8 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
timeout 300 python tools/march_variants.py 7 50 $gens > gpurun_out/${tag}_variants.jsonl 2> gpurun_out/${tag}_variants.err
cat gpurun_out/${tag}_variants.jsonl; tail -3 gpurun_out/${tag}_variants.err
if [ "$ncugen" != "none" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 3 -c 1 -o gpurun_out/${tag}_${ncugen} -f python tools/march_variants.py 7 3 $ncugen > gpurun_out/${tag}_ncu.log 2>&1
  tail -2 gpurun_out/${tag}_ncu.log
fi
