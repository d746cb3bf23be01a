NG=${NG:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555"
mkdir -p gpurun_out/s2h
timeout 300 $TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nvlink --no-cpu-baseline --no-variants > gpurun_out/s2h/n${NG}_nvlink.json 2>> gpurun_out/s2h/n$NG.err
timeout 300 $TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nccl --no-cpu-baseline --no-variants > gpurun_out/s2h/n${NG}_nccl.json 2>> gpurun_out/s2h/n$NG.err
for f in gpurun_out/s2h/n${NG}_*.json; do echo $f; python -c "
import json,sys
j=json.loads(open('$f').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['exposed_exchange_us_per_step'], j['e2e']['value'], j['config']['grad_exchange'][:100])"; done
