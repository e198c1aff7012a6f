# usage: bash tools/n2_async.sh N  -- multi-GPU parity check, then the vmult bench with the asynchronous / synchronous peer-memory import and NCCL
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_multi_gpu.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|^{\"world" | tail -n 6
run() { timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 3 --no-cpu --no-e2e --no-solve --no-step 2>gpurun_out/async_$1.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'vmult ms', round(d['ms_per_step'],4), 'GDoF/s', round(d['value']/1e9,1), 'check', d['marching_vs_generic_rel_l2'], 'parity', d['multi_gpu_parity'])
"; tail -n 2 gpurun_out/async_$1.err | grep -i "error\|Traceback" ; }
run async
DGX_P2P_ASYNC=0 run sync
