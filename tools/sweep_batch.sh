#!/bin/bash
# batch-size sweep of the device-resident path
for cfg in "248 16" "248 28" "248 56" "248 84" "504 13" "504 26" "504 39" "120 64" "120 201" "1016 7"; do
  set -- $cfg
  python bench.py --steps 8 --warmup 3 --cs $1 --batch $2 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('cs',$1,'batch',$2,'value %.1f MP/s  e2e %.1f  ms/step %.2f  conv TF %.0f  sm_mhz %s %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['achieved'],d['clocks']['sm_mhz'],d['clocks']['reasons']))"
done
