#!/bin/bash
# samples SM / memory clocks, power and throttle reasons every 50 ms while "$@" runs; prints min/median/max
out=$(mktemp)
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 50 > "$out" &
pid=$!
"$@"
kill $pid
python - "$out" <<'PY'
import sys, numpy as np
rows=[l.strip().split(', ') for l in open(sys.argv[1]) if l.strip()]
a=np.array([[float(r[0]),float(r[1]),float(r[2]),float(r[3])] for r in rows if len(r)>=5])
cap=sum(1 for r in rows if len(r)>=5 and r[4].startswith('Active'))
busy=a[a[:,2]>400]
for nm,b in (('all',a),('busy(>400W)',busy)):
    if len(b): print(nm,'samples',len(b),'sm',b[:,0].min(),np.median(b[:,0]),b[:,0].max(),'mem',b[:,1].min(),np.median(b[:,1]),b[:,1].max(),'power',b[:,2].min(),np.median(b[:,2]),b[:,2].max(),'limit',b[0,3],'cap_samples',cap)
PY
