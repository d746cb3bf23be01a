NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555"
timeout 300 $TR bench.py --gpus $NG --check 2> gpurun_out/r2_check_n$NG.err | tee gpurun_out/r2_check_n$NG.txt
timeout 300 $TR tools/bench_xchg.py 2> gpurun_out/r2_xchg_n$NG.err | tee gpurun_out/r2_xchg_n$NG.txt
timeout 300 $TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nvlink > gpurun_out/r2_n${NG}_nvlink.json 2>> gpurun_out/r2_n$NG.err
SIG_SYNC_CHUNKS=2 timeout 300 $TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nvlink > gpurun_out/r2_n${NG}_nvlink_2pieces.json 2>> gpurun_out/r2_n$NG.err
timeout 300 $TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nccl > gpurun_out/r2_n${NG}_nccl.json 2>> gpurun_out/r2_n$NG.err
for f in gpurun_out/r2_n${NG}_*.json; do echo $f; python -c "
import json,sys
j=json.loads(open('$f').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['exposed_exchange_us_per_step'], j['e2e']['value'], j['config']['grad_exchange'][:150])"; done
