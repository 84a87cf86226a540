set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_va_models.py -q -m gpu -k "dff_adaptive" -s 2>&1 | tail -15 > gpurun_out/r02m_dff.log
python bench.py > gpurun_out/r02m_c3.json 2> gpurun_out/r02m_c3.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_c3_s2.json 2> gpurun_out/r02m_c3_s2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_ncu1.log 2>&1
CB200_SEGMENTS=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_c3_seg1.json 2> gpurun_out/r02m_c3_seg1.err && \
CB200_SEGMENTS=1 ncu --set full --clock-control none --import-source on -k regex:cb200_spec_tran_fixed_kernel -s 3 -c 1 -o gpurun_out/r02m_c3_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_ncu2.log 2>&1
ncu -i gpurun_out/r02m_c3_full.ncu-rep --page raw --csv > gpurun_out/r02m_c3_spec_tran_fixed_ncu_full.csv 2>/dev/null
python bench.py --workload c4 --steps 2 > gpurun_out/r02m_c4.json 2> gpurun_out/r02m_c4.err
python bench.py --workload c1 --steps 3 > gpurun_out/r02m_c1.json 2> gpurun_out/r02m_c1.err
python bench.py --workload c2 --steps 5 > gpurun_out/r02m_c2.json 2> gpurun_out/r02m_c2.err
ls -la gpurun_out
