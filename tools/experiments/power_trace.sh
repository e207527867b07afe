#!/bin/bash
# power / clock trace during a long bench run (kills the sampler by PID)
nvidia-smi --query-gpu=clocks.sm,power.draw.instant,power.draw.average --format=csv,noheader,nounits -lms 20 > gpurun_out/pw.csv &
SMI=$!
sleep 1
timeout 200 python bench.py --steps 100 --no-cpu-baseline | cut -c1-160
kill $SMI
python - <<'PY'
rows=[l.strip().split(', ') for l in open('gpurun_out/pw.csv') if l.strip()]
rows=[(float(a),float(b),float(c)) for a,b,c in rows]
busy=[r for r in rows if r[1]>400]
print("samples %d busy %d" % (len(rows), len(busy)))
if busy:
    import statistics as st
    print("busy: sm MHz median %.0f min %.0f | instant W median %.0f max %.0f | avg W median %.0f max %.0f" % (st.median(r[0] for r in busy), min(r[0] for r in busy), st.median(r[1] for r in busy), max(r[1] for r in busy), st.median(r[2] for r in busy), max(r[2] for r in busy)))
PY
