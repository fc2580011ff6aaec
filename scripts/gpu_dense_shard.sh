#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/prof_dense_shard.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_dense_shard.csv python scripts/prof_dense_shard.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_dense_shard.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
ik=rows[hdr].index('Kernel Name'); iv=rows[hdr].index('Metric Value')
sel=[(r[ik].split('(')[0][-40:], float(r[iv].replace(',',''))/1e6) for r in rows[hdr+1:] if len(r)>iv and ('vs::' in r[ik])]
# last call = last 12 or so launches
for name,ms in sel[-14:]: print(f"{ms:8.3f} ms  {name}")
PY
