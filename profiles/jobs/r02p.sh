# the ONE C3 sweep partitioned over 2 GPUs (strong scaling), launched as the driver launches it
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02p_c3_2gpu.json 2> gpurun_out/r02p_c3_2gpu.err
tail -c 600 gpurun_out/r02p_c3_2gpu.err; head -c 600 gpurun_out/r02p_c3_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02p_ref_2gpu.json 2> gpurun_out/r02p_ref_2gpu.err
head -c 400 gpurun_out/r02p_ref_2gpu.json
