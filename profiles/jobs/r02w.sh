# PSP103 on the GPU: the ring's parity test and the ring bench line (100 ns of the 1 us benchmark span, BDF controller)
cd $GRAFT_REPO_ROOT
timeout 240 python -m pytest tests/test_va_models.py -q -m gpu -k "psp103_ring or bsim4" -s 2>&1 | tail -8 > gpurun_out/r02w_psp_test.log
tail -5 gpurun_out/r02w_psp_test.log
CB200_RING_TSTOP=1e-7 CB200_RING_MAXPOINTS=8192 timeout 420 python bench.py --workload ring --steps 2 --warmup 3 > gpurun_out/r02w_ring.json 2> gpurun_out/r02w_ring.err
tail -c 500 gpurun_out/r02w_ring.err; head -c 1500 gpurun_out/r02w_ring.json
