python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
cat > /tmp/show.py <<'PY'
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); n=d['steps']; print(d['config']['workload'][:30], d['n_gpus'], 'value %.4g e2e %.4g ms/step %.2f waves %.1f reach %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['waves_per_step'], d['config'].get('plans_reaching_goal')), {k: round(v/n,2) for k,v in d['stage_ms'].items()}, flush=True)
    elif 'rror' in l or 'Traceback' in l: print(l.strip())
PY
python bench.py --steps 10 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_r1_final.json 2>gpurun_out/bench_r1_final.err; python /tmp/show.py < gpurun_out/bench_r1_final.json; tail -2 gpurun_out/bench_r1_final.err
python bench.py --steps 5 --warmup 2 --no-cpu --wave 512 2>&1 | python /tmp/show.py
python bench.py --steps 5 --warmup 2 --no-cpu --config c3 2>&1 | python /tmp/show.py
python bench.py --steps 5 --warmup 2 --no-cpu --config c2 2>&1 | python /tmp/show.py
