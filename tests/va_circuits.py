"""Circuits built from the reference's VADistiller models (sp_mos1, sp_diode): the builders live in
the package (cadnip_b200.workloads) so that bench.py and build() do not import the test tree;
re-exported here for the fixture generator and the CPU tests."""
from cadnip_b200.workloads import (VA_DIR, DFF_DIR, FIXTURES, C3_PWL_T, C3_PWL_Y, mos1_corner, diode_chain,
                                   diode_rs_cap, mos1_inverter, mos1_c3, c3_lane_exprs, mos1_dff, mos1_ring,
                                   lower_fixture, load_fixture, load_workload, fixture_path)
