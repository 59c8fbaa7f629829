# multi-GPU check on one box: parity tests (1 GPU), bench at N=1 and N=$1 (torchrun, NCCL)
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; cut -c1-400 gpurun_out/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_ref_n$N.json
ls -la gpurun_out
