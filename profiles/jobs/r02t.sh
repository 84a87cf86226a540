# final kernels of the round: bypass with the executed-step counter.  Parity test, C3 line, launch list and a full
# ncu capture of one whole-sweep launch of the dominant kernel.
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_va_models.py -q -m gpu -k "pair_mode" -s 2>&1 | tail -6 > gpurun_out/r02t_tests.log
tail -4 gpurun_out/r02t_tests.log
python bench.py > gpurun_out/r02t_c3.json 2> gpurun_out/r02t_c3.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_c3_s2.json 2> gpurun_out/r02t_c3_s2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02t_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_ncu1.log 2>&1
CB200_SEGMENTS=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_c3_seg1.json 2> gpurun_out/r02t_c3_seg1.err && \
CB200_SEGMENTS=1 ncu --set full --clock-control none --import-source on -k regex:cb200_spec_tran_fixed_kernel -s 3 -c 1 -o gpurun_out/r02t_c3_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02t_ncu2.log 2>&1
ncu -i gpurun_out/r02t_c3_full.ncu-rep --page raw --csv > gpurun_out/r02t_c3_spec_tran_fixed_ncu_full.csv 2>/dev/null
python bench.py --workload c2 > gpurun_out/r02t_c2.json 2> gpurun_out/r02t_c2.err
python bench.py --workload c1 --steps 3 > gpurun_out/r02t_c1.json 2> gpurun_out/r02t_c1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02t_c*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]; print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["tran_kernel_ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", r.get("frac"), "exec", r.get("lane_steps_executed"), "/", r.get("lane_steps_total"), d.get("parity", {}).get("max_abs_diff_vs_oracle"))
    except Exception as e:
        print(f, "ERR", e)
PY
