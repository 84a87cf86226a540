# the ring bench line on the first 20 ns of the benchmark span (1024 supply-voltage lanes, BDF controller)
cd $GRAFT_REPO_ROOT
CB200_RING_TSTOP=2e-8 CB200_RING_MAXPOINTS=2048 timeout 200 python bench.py --workload ring --steps 1 --warmup 3 > gpurun_out/r02z_ring.json 2> gpurun_out/r02z_ring.err
tail -c 600 gpurun_out/r02z_ring.err; head -c 2500 gpurun_out/r02z_ring.json
