mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/r4p_gpus_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(timeout 600 $TR tests/peer_worker.py 2>&1 | tail -25) > gpurun_out/r4p_peer_n$N.log
(timeout 900 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 2> gpurun_out/r4p_ref_n$N.err | tail -1) > gpurun_out/r4p_ref_n$N.json
(timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 2> gpurun_out/r4p_bench_n$N.err | tail -1) > gpurun_out/r4p_bench_n$N.json
(timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 --gather nccl 2> gpurun_out/r4p_bench_nccl_n$N.err | tail -1) > gpurun_out/r4p_bench_nccl_n$N.json
tail -5 gpurun_out/r4p_peer_n$N.log; cut -c1-400 gpurun_out/r4p_ref_n$N.json; cut -c1-1200 gpurun_out/r4p_bench_n$N.json; tail -3 gpurun_out/r4p_bench_n$N.err; cut -c1-200 gpurun_out/r4p_bench_nccl_n$N.json
