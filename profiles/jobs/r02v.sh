# final state of the round: the full GPU suite, smoke(), the default bench line and the reference arm
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu 2>&1 | grep -v "^$" | tail -15 > gpurun_out/r02v_suite.log
tail -4 gpurun_out/r02v_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02v_smoke.log 2>&1; tail -3 gpurun_out/r02v_smoke.log
python bench.py > gpurun_out/r02v_c3.json 2> gpurun_out/r02v_c3.err; head -c 400 gpurun_out/r02v_c3.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02v_reference.json 2> gpurun_out/r02v_reference.err; head -c 300 gpurun_out/r02v_reference.json; echo
