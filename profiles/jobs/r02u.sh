# experiment: block-lockstep time loop (CB200_LOCKSTEP=1) against the default, C3
cd $GRAFT_REPO_ROOT
CB200_LOCKSTEP=1 python bench.py > gpurun_out/r02u_c3_lock256.json 2> gpurun_out/r02u_c3_lock256.err
CB200_LOCKSTEP=1 CB200_SPEC_BLOCK=128 python bench.py --no-cpu-baseline > gpurun_out/r02u_c3_lock128.json 2> gpurun_out/r02u_c3_lock128.err
CB200_LOCKSTEP=1 CB200_SPEC_BLOCK=64 python bench.py --no-cpu-baseline > gpurun_out/r02u_c3_lock64.json 2> gpurun_out/r02u_c3_lock64.err
CB200_LOCKSTEP=1 python bench.py --lanes 12500 --no-cpu-baseline > gpurun_out/r02u_c3_lock256_12500.json 2> gpurun_out/r02u_c3_lock256_12500.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02u_c*.json")):
    try:
        d = json.load(open(f)); r = d["roofline"]; print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "kernel", round(d["tran_kernel_ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", r.get("frac"), "exec", r.get("lane_steps_executed"), "/", r.get("lane_steps_total"), d.get("parity", {}).get("max_abs_diff_vs_oracle"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-600:])
PY
