cd $GRAFT_REPO_ROOT
timeout 125 python -m pytest tests/test_va_models.py -q -m gpu -k "full_size_properties_c3" -s 2>&1 | tail -6 | cut -c1-500 > gpurun_out/r02zz_c3_full_test.log
cat gpurun_out/r02zz_c3_full_test.log
