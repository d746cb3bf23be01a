# usage: bash tools/sweep_env.sh VAR v1 v2 ...   -> graph-replay step time of bench.py for each value
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', j['value'], j['ms_per_step'])"
done
