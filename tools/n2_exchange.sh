NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555"
$TR bench.py --gpus $NG --check 2> gpurun_out/r2_check.err | tee gpurun_out/r2_check_n$NG.txt
$TR tools/bench_xchg.py 2> gpurun_out/r2_xchg.err | tee gpurun_out/r2_xchg_n$NG.txt
$TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nvlink > gpurun_out/r2_n${NG}_nvlink.json 2>> gpurun_out/r2_n2.err
$TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nccl > gpurun_out/r2_n${NG}_nccl.json 2>> gpurun_out/r2_n2.err
SIG_XCHG_MULTICAST=0 $TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nvlink > gpurun_out/r2_n${NG}_nvlink_p2p.json 2>> gpurun_out/r2_n2.err
SIG_SYNC_CHUNKS=2 $TR bench.py --gpus $NG --steps 100 --warmup 5 --exchange nvlink > gpurun_out/r2_n${NG}_nvlink_2pieces.json 2>> gpurun_out/r2_n2.err
for f in gpurun_out/r2_n${NG}_*.json; do echo $f; python -c "
import json,sys
j=json.loads(open('$f').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['exposed_exchange_us_per_step'], j['e2e']['value'], j['config']['grad_exchange'][:150])"; done
