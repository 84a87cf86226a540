"""Writes tests/golden/reference_kat.json.

The reference (NyanCAD/Cadnip.jl) is Julia and cannot be executed in the build
container, so these golden vectors are the known-answer constants of the
reference's OWN tests, transcribed by hand; every entry cites the test file:line it
comes from (paths relative to the reference repository).  Nothing here was computed
by this repository's code.  Re-run to regenerate the JSON.
"""
import json
import math
import os

vt = 0.026
kat = {
    "pnjlim": [  # test/mna/pcnr.jl:114-146 ; args (vnew, vold, vt, vcrit) -> (vlim, limited)
        {"args": [0.5, 0.49, 0.026, 0.7], "vlim": 0.5, "limited": False, "exact": True, "src": "test/mna/pcnr.jl:116"},
        {"args": [-0.3, -0.3, 0.026, 0.6588], "vlim": -0.3, "limited": False, "src": "test/mna/pcnr.jl:120-124"},
        {"args": [0.0, 0.0, 0.026, 0.6588], "vlim": 0.0, "limited": False, "src": "test/mna/pcnr.jl:120-124"},
        {"args": [0.3, 0.3, 0.026, 0.6588], "vlim": 0.3, "limited": False, "src": "test/mna/pcnr.jl:120-124"},
        {"args": [0.6588, 0.6588, 0.026, 0.6588], "vlim": 0.6588, "limited": False, "src": "test/mna/pcnr.jl:120-124"},
        {"args": [5.0, 0.0, 0.026, 0.66], "vlim": 0.026 * math.log(5.0 / 0.026), "limited": True, "src": "test/mna/pcnr.jl:135-137"},
        {"args": [-10.0, 0.5, 0.026, 0.66], "vlim": -1.5, "limited": True, "exact": True, "src": "test/mna/pcnr.jl:140-141"},
        {"args": [-0.5, 0.0, 0.026, 0.66], "vlim": -0.5, "limited": False, "exact": True, "src": "test/mna/pcnr.jl:144-145"},
    ],
    "pnjlim_compression": {  # test/mna/pcnr.jl:128-132: vold < vlim < vnew and vlim < 1.0
        "args": [5.0, 0.6, 0.026, 0.6588], "src": "test/mna/pcnr.jl:128-132"},
    "diode_iv": {  # test/mna/pcnr.jl:152-172
        "Is": 1e-14, "nVt": 0.026,
        "exact_at": 0.7,
        "I_exact": 1e-14 * (math.exp(0.7 / 0.026) - 1.0),
        "G_exact": 1e-14 / 0.026 * math.exp(0.7 / 0.026),
        "src": "test/mna/pcnr.jl:152-172"},
    "coo_to_csc": [  # test/mna/precompile.jl:18-68
        {"I": [1, 2, 1, 3, 2], "J": [1, 1, 2, 2, 3], "V": [1.0, 2.0, 3.0, 4.0, 5.0], "n": 3,
         "src": "test/mna/precompile.jl:18-41"},
        {"I": [1, 1, 2], "J": [1, 1, 2], "V": [1.0, 2.0, 3.0], "n": 2, "dup": [0, 1],
         "dense": [[3.0, 0.0], [0.0, 3.0]], "src": "test/mna/precompile.jl:43-68"},
    ],
    "rectifier": {  # test/mna/pcnr.jl:268-285, :330-343 ; test/mna/precompile.jl:205-242
        "n": 4, "n_limits": 1, "out_range": [0.55, 0.75], "lim_vs_nolim_atol": 1e-6,
        "pcnr_iters_max": 10, "src": "test/mna/pcnr.jl:268-285,330-343"},
    "chain": {  # test/mna/pcnr.jl:291-324, :345-350
        "n_limits": 3, "vd_range": [0.6, 0.85], "equal_rtol": 1e-2, "pcnr_iters_max": 10,
        "src": "test/mna/pcnr.jl:291-324,345-350"},
    "dc_linear": [  # test/mna/core.jl:509-637
        {"name": "divider", "expect": {"vcc": 5.0, "out": 2.5}, "atol": 1e-10, "src": "test/mna/core.jl:509-520"},
        {"name": "divider_unequal", "expect": {"out": 5.0 / 3.0}, "atol": 1e-10, "src": "test/mna/core.jl:522-532"},
        {"name": "isrc_resistor", "expect": {"n1": 1.0}, "atol": 1e-10, "src": "test/mna/core.jl:534-543"},
        {"name": "two_vsources", "expect": {"vcc": 5.0, "mid": 3.0}, "atol": 1e-10, "src": "test/mna/core.jl:545-557"},
        {"name": "vccs_amp", "expect": {"inp": 1.0, "out": 10.0}, "atol": 1e-10, "src": "test/mna/core.jl:559-572"},
        {"name": "vcvs_inv", "expect": {"inp": 0.5, "out": -5.0}, "atol": 1e-10, "src": "test/mna/core.jl:574-584"},
        {"name": "ccvs_transres", "expect": {"out": 1.0}, "atol": 1e-6, "src": "test/mna/core.jl:586-599"},
        {"name": "cccs_mirror", "expect": {"out": 2.0}, "atol": 1e-10, "src": "test/mna/core.jl:601-615"},
        {"name": "multinode", "expect": {"center": 9.0 / (1.0 + 0.5 + 1.0 / 3.0)}, "atol": 1e-10, "src": "test/mna/core.jl:617-637"},
    ],
    "rc_charge": {  # test/mna/core.jl:785-857
        "Vcc": 5.0, "R": 1000.0, "C": 1e-6, "times_tau": [0.0, 0.5, 1.0, 2.0, 3.0, 5.0], "rtol": 1e-3,
        "src": "test/mna/core.jl:785-857,859-912"},
    "sweep_divider": {  # test/sweep.jl:299-313 : I_V = -1/(R1+R2), atol 1e-8 (deftol)
        "R1": [100.0, 100.0, 2000.0], "R2": [100.0, 100.0, 2000.0], "atol": 1e-8, "src": "test/sweep.jl:299-313"},
    "pulse": {  # test/transients.jl:112-175: plateau values across periods; defaults codegen.jl:2697-2703
        "v1": 0.0, "v2": 5.0, "td": 1e-3, "tr": 1e-4, "tf": 1e-4, "pw": 1e-3, "per": 4e-3,
        "points": [[0.0, 0.0], [0.5e-3, 0.0], [1.05e-3, 2.5], [1.5e-3, 5.0], [2.15e-3, 2.5], [3.0e-3, 0.0],
                   [5.5e-3, 5.0], [9.5e-3, 5.0], [7.0e-3, 0.0]],
        "src": "src/mna/devices.jl:85-103 semantics as exercised by test/transients.jl:112-175"},
    "breakpoints_pulse": {  # test/mna/breakpoints.jl:22-60
        "wave": [0.0, 1.0, 1e-6, 1e-6, 1e-6, 1e-3, 2e-3],
        "edges": [1e-6, 2e-6, 1.002e-3, 1.003e-3], "period": 2e-3,
        "src": "test/mna/breakpoints.jl:22-40"},
}
here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "reference_kat.json"), "w") as f:
    json.dump(kat, f, indent=1)
print("wrote reference_kat.json")
