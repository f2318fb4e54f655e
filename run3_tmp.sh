cat > /tmp/show.py <<'PY'
import json,sys,os
tag=os.environ.get('TAG','')
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); n=d['steps']; d['exchange']=tag
        print(tag, d['config']['workload'][:24], d['n_gpus'], d['scaling'], 'value %.4g e2e %.4g ms/step %.2f waves %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['waves_per_step']), {k: round(v/n,2) for k,v in d['stage_ms'].items()}, flush=True)
        open('gpurun_out/scale_r1.jsonl','a').write(json.dumps(d)+'\n')
    elif 'rror' in l or 'Traceback' in l: print(l.strip())
PY
rm -f gpurun_out/scale_r1.jsonl
timeout 300 python -m pytest tests/test_gpu_multi.py -q --timeout 300 2>&1 | tail -2
TAG=single timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | python /tmp/show.py
for ex in peer nccl; do
  for n in 2 4 8; do
    TAG=$ex LQRRT_EXCHANGE=$ex timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --no-cpu 2>&1 | python /tmp/show.py
  done
done
TAG=peer LQRRT_EXCHANGE=peer timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --scaling weak 2>&1 | python /tmp/show.py
echo "== c5k"
TAG=single timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu --config c5k 2>&1 | python /tmp/show.py
TAG=peer timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu --config c5k --scaling weak 2>&1 | python /tmp/show.py
TAG=peer timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 1 --no-cpu --config c5k --scaling strong 2>&1 | python /tmp/show.py
