run() { env "$@" python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', j['value'], j['ms_per_step'])"; }
run SIG_ALIGN_SMS=0
run SIG_ALIGN_SMS=116
run SIG_ALIGN_SMS=100
run SIG_ALIGN_WAVES=2
run SIG_ALIGN_WAVES=4
run SIG_ALIGN_WAVES=2 SIG_PRIO=1
run SIG_ALIGN_WAVES=4 SIG_PRIO=1
run SIG_ALIGN_WAVES=8 SIG_PRIO=1
run SIG_ALIGN_SMS=116 SIG_ALIGN_WAVES=4 SIG_PRIO=1
