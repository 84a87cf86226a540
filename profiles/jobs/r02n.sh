# BDF controller on the device: parity vs the oracle, and warp / block mappings vs lane-per-thread
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bdf or lane_per_warp or adaptive" -s 2>&1 | tail -60 > gpurun_out/r02n_bdf_tests.log
cat gpurun_out/r02n_bdf_tests.log | tail -30
