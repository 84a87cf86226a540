# full GPU suite + smoke on the final kernel sources (BDF controller, restart-step fix)
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/r02o_suite.log
tail -5 gpurun_out/r02o_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke.log 2>&1; tail -3 gpurun_out/r02o_smoke.log
python bench.py > gpurun_out/r02o_c3.json 2> gpurun_out/r02o_c3.err
python bench.py --workload c4 --steps 2 > gpurun_out/r02o_c4.json 2> gpurun_out/r02o_c4.err
CB200_ADAPTIVE_METHOD=bdf python bench.py --workload c4 --steps 2 > gpurun_out/r02o_c4_bdf.json 2> gpurun_out/r02o_c4_bdf.err
