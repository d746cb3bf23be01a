run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-variants "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', j['value'], j['ms_per_step'])"; }
run --only sim
run --only align
python tools/timeline.py sim 2>/dev/null
python tools/timeline.py align 2>/dev/null
