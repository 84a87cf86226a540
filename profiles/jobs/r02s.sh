# quiescent-step bypass of the specialised fixed-step kernel: parity (three builds bit-identical, vs table-driven,
# vs oracle) and the C3 / C2 lines with and without it
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_va_models.py tests/test_gpu_parity.py -q -m gpu -k "pair_mode or specialised or spec or c3 or full_size" -s 2>&1 | tail -25 > gpurun_out/r02s_tests.log
tail -8 gpurun_out/r02s_tests.log
python bench.py > gpurun_out/r02s_c3.json 2> gpurun_out/r02s_c3.err
CB200_NO_BYPASS=1 python bench.py --no-cpu-baseline > gpurun_out/r02s_c3_plain.json 2> gpurun_out/r02s_c3_plain.err
python bench.py --lanes 12500 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_c3_12500.json 2> gpurun_out/r02s_c3_12500.err
python bench.py --lanes 25000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_c3_25000.json 2> gpurun_out/r02s_c3_25000.err
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r02s_c2.json 2> gpurun_out/r02s_c2.err
python bench.py --workload c1 --steps 3 --no-cpu-baseline > gpurun_out/r02s_c1.json 2> gpurun_out/r02s_c1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02s_c*.json")):
    try:
        d = json.load(open(f)); print(f, d["config"].get("lanes_total"), "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["tran_kernel_ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"].get("frac"), d.get("parity", {}).get("max_abs_diff_vs_oracle"))
    except Exception as e:
        print(f, "ERR", e)
PY
