#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x > $OUT/r2e_pytest_train.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2e_pytest_train.log
tail -5 $OUT/r2e_pytest_train.log
timeout 900 python bench.py --workload ppo --ppo-envs 8192 --ppo-clips 8 --steps 2 --warmup 1 > $OUT/r2e_bench_ppo_small.json 2> $OUT/r2e_bench_ppo_small.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$OUT/r2e_bench_ppo_small.json')); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['learner'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r2e_launches_ppo.csv python bench.py --workload ppo --ppo-envs 8192 --ppo-clips 8 --steps 1 --warmup 1 > $OUT/r2e_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2e_launches_ppo.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    name=r[4].split('(')[0][:70]; 
    try: v=float(r[-1].replace(',',''))
    except: continue
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
print('launches',len(rows),'total us',tot/1e3)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:28]: print(f'{v[1]/1e3:10.1f} us {v[0]:6d}  {k}')
PY
