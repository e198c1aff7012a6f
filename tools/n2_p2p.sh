# usage: bash tools/n2_p2p.sh N   -- multi-GPU parity check, then the bench with and without the peer-memory ghost import
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_multi_gpu.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -n 12
run() { timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'vmult ms', round(d['ms_per_step'],4), 'mg', round(d['mg_solve']['ms_per_solve'],2), round(d['mg_solve']['ms_per_solve_cached_estimates'],2), 'its', d['mg_solve']['cg_iterations'], 'step', round(d['trbdf2_step']['ms_per_step'],2), 'parity', d['multi_gpu_parity'])
"; }
run p2p
DGX_P2P=0 run nccl
