NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29556"
SIG_EXCHANGE=nvlink $TR tools/timeline.py 2>/dev/null > gpurun_out/r2_timeline_n${NG}_nvlink.txt
SIG_EXCHANGE=nvlink SIG_SYNC_CHUNKS=2 $TR tools/timeline.py 2>/dev/null > gpurun_out/r2_timeline_n${NG}_nvlink_2p.txt
SIG_EXCHANGE=nccl $TR tools/timeline.py 2>/dev/null > gpurun_out/r2_timeline_n${NG}_nccl.txt
tail -n 40 gpurun_out/r2_timeline_n${NG}_nvlink.txt
