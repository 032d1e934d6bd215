#!/bin/bash
# power of each store pattern: nvidia-smi samples (100 ms) while the pattern loops for 2 s; the mean of the samples after the first second is reported
for spec in "S 768" "A 768" "D4 768" "D8 768" "C 768" "B 768" "A 256" "S 256"; do
  set -- $spec
  nvidia-smi --query-gpu=power.draw.instant,clocks.sm --format=csv,noheader,nounits -lms 100 > /tmp/pw.txt &
  SMI=$!
  sleep 0.3
  tools/ubench/store_patterns $1 $2 2.0
  kill $SMI; wait $SMI 2>/dev/null
  python3 - <<'PY'
rows=[l.strip().split(',') for l in open('/tmp/pw.txt') if l.strip()]
v=[(float(a),float(b)) for a,b in rows]
h=v[len(v)//2:-1] or v
print("    power mean(2nd half) %.0f W  sm %.0f MHz  (%d samples)" % (sum(x[0] for x in h)/len(h), sum(x[1] for x in h)/len(h), len(h)))
PY
done
