# usage: bash tools/bench_n.sh N [extra bench args]  -> gpurun_out/r2_bench_nN.json
N=${1:-2}; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 "$@" > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - <<PY
import json
for l in open('gpurun_out/r2_bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=$N vmult ms', round(d['ms_per_step'],4), 'GDoF/s', round(d['value']/1e9,1), 'mg', round(d['mg_solve']['ms_per_solve'],2), round(d['mg_solve']['ms_per_solve_cached_estimates'],2), 'its', d['mg_solve']['cg_iterations'], 'levels', d['mg_solve']['levels'], 'launches', d['mg_solve']['gpu_launches'], 'step', round(d['trbdf2_step']['ms_per_step'],2) if d.get('trbdf2_step') else None, 'e2e', d['e2e']['ms_per_step'] if d.get('e2e') else None, 'parity', d['multi_gpu_parity'])
PY
tail -n 3 gpurun_out/r2_bench_n$N.err
