# PSP103 ring on the GPU, second attempt within the remaining budget: parity test, and the bench line on the first
# 20 ns of the benchmark span
cd $GRAFT_REPO_ROOT
timeout 120 python -m pytest tests/test_va_models.py -q -m gpu -k "psp103_ring" -s 2>&1 | tail -6 > gpurun_out/r02x_psp_test.log
tail -4 gpurun_out/r02x_psp_test.log
CB200_RING_TSTOP=2e-8 CB200_RING_MAXPOINTS=2048 timeout 240 python bench.py --workload ring --steps 1 --warmup 3 > gpurun_out/r02x_ring.json 2> gpurun_out/r02x_ring.err
tail -c 400 gpurun_out/r02x_ring.err; head -c 1200 gpurun_out/r02x_ring.json
