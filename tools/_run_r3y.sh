mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for c in 5 4; do
(timeout 600 $TR bench.py --config $c --gpus $N --steps 4 --warmup 3 --no-cpu-baseline 2> gpurun_out/r3y_bench_c${c}_n$N.err | tail -1) > gpurun_out/r3y_bench_c${c}_n$N.json
cut -c1-220 gpurun_out/r3y_bench_c${c}_n$N.json; tail -2 gpurun_out/r3y_bench_c${c}_n$N.err
done
