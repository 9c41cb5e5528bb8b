# 2-GPU check: gpurun --gpus 2 --timeout 900 -- 'bash tools/scale2_run.sh TAG'
T=${1:-r2}
python -m pytest tests/test_multi_gpu.py -m gpu -q -rs 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_scale_2.json 2> gpurun_out/${T}_scale_2.err
echo rc=$?; tail -3 gpurun_out/${T}_scale_2.err; cut -c1-600 gpurun_out/${T}_scale_2.json
