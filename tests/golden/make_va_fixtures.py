"""Generates tests/golden/va_<name>.json.gz: circuits built from the reference's VADistiller
Verilog-A models (models/VADistillerModels.jl/va/{mos1,diode}.va), lowered by the product's
emitter HERE (where /root/reference is mounted).  A fixture is the LoweredCircuit as gzip'd JSON
(LoweredCircuit.save: integer / float tables, the per-lane parameters, and the emitted CUDA header
/ oracle C text) -- reviewable with zcat, and loading it executes nothing -- so the GPU tests can
run sp_mos1 / sp_diode circuits on a box that has no reference tree.

    python tests/golden/make_va_fixtures.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import va_circuits  # noqa: E402


def main():
    models = {}
    for name in va_circuits.FIXTURES:
        lc = va_circuits.lower_fixture(name, models)
        path = va_circuits.fixture_path(name)     # parsed models do not travel; their emitted text does
        lc.save(path)
        print(f"{name}: n={lc.n} (nodes {lc.n_nodes}, currents {lc.n_currents}, charges {lc.n_charges}, "
              f"limits {lc.n_limits}), P={lc.P}, header {len(lc.va_cuda_header) // 1024} KiB -> "
              f"{os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    main()
