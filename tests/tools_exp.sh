for cfg in "TTS_STAGGER=0" "TTS_STAGGER=25" "TTS_GROUP_ROWS=8" "TTS_DECODE_V2=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line); print(d['value'], d['roofline']['us_per_decode_step'], d['roofline']['frac'])"
done
