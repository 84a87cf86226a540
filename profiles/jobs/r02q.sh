# strong scaling of the ONE C3 sweep over 4 and 8 GPUs, launched as the driver launches it
cd $GRAFT_REPO_ROOT
for N in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_c3_${N}gpu.json 2> gpurun_out/r02q_c3_${N}gpu.err
tail -c 300 gpurun_out/r02q_c3_${N}gpu.err; head -c 300 gpurun_out/r02q_c3_${N}gpu.json; echo
done
