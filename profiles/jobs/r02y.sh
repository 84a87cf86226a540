# the PSP103 ring parity test with its final bar
cd $GRAFT_REPO_ROOT
timeout 150 python -m pytest tests/test_va_models.py -q -m gpu -k "psp103_ring" -s 2>&1 | tail -12 | cut -c1-600 > gpurun_out/r02y_psp_test.log
cat gpurun_out/r02y_psp_test.log
