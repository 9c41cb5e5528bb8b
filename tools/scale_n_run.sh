# N-GPU check: gpurun --gpus N --timeout 900 -- 'bash tools/scale_n_run.sh N TAG'
N=${1:-8}; T=${2:-r2}
python -m pytest tests/test_multi_gpu.py -m gpu -q -rs 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_scale_$N.json 2> gpurun_out/${T}_scale_$N.err
echo rc=$?; tail -3 gpurun_out/${T}_scale_$N.err; cut -c1-400 gpurun_out/${T}_scale_$N.json
