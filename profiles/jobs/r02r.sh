# pair mode (two warps per 32 lanes) of the specialised fixed-step kernel: parity and the small-sweep regime
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_va_models.py tests/test_gpu_parity.py -q -m gpu -k "pair_mode or bdf or specialised or spec" -s 2>&1 | tail -25 > gpurun_out/r02r_tests.log
tail -8 gpurun_out/r02r_tests.log
for L in 3125 6250 12500 25000; do
CB200_PAIR=0 python bench.py --lanes $L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_c3_${L}_single.json 2> gpurun_out/r02r_c3_${L}_single.err
python bench.py --lanes $L --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_c3_${L}_auto.json 2> gpurun_out/r02r_c3_${L}_auto.err
done
CB200_PAIR=1 python bench.py --lanes 25000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_c3_25000_pair.json 2> gpurun_out/r02r_c3_25000_pair.err
python bench.py --workload c1 --steps 3 > gpurun_out/r02r_c1.json 2> gpurun_out/r02r_c1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02r_c3_*.json")) + ["gpurun_out/r02r_c1.json"]:
    try:
        d = json.load(open(f)); print(f, d["config"].get("lanes_total"), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["tran_kernel_ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), d.get("parity", {}).get("max_abs_diff_vs_oracle"))
    except Exception as e:
        print(f, "ERR", e)
PY
