"""cadnip-b200: B200-native batched MNA Newton/transient hot path behind Cadnip's
MNACircuit / dc! / tran! / CircuitSweep API.

The directory is named ``cadnip.jl_b200`` (not importable as written); import it as
``cadnip_b200`` through the shim module at the repository root.

Host mirror of the reference interface (``!`` dropped from Julia names):
``MNAContext, get_node, stamp, Resistor ... SimpleMOSFET, MNASpec, MNACircuit,
alter, Sweep, ProductSweep, TandemSweep, SerialSweep, CircuitSweep, SweepResult,
dc, tran``.  Everything numerical runs in ``csrc/libcadnip_b200.so`` (CUDA,
sm_100a) through the C ABI of ``include/cadnip_b200.h``; there is no CPU fallback.
"""
from .mna import (MNAContext, ZERO_VECTOR, ZeroVector, CurrentIndex, ChargeIndex, LimitIndex,
                  get_node, alloc_current, alloc_internal_node, alloc_limit, resolve_index,
                  stamp_G, stamp_C, stamp_b, system_size, reset_for_restamping, stamp,
                  Resistor, Capacitor, Inductor, VoltageSource, CurrentSource, VCVS, VCCS, CCVS,
                  CCCS, Diode, DiodeWithCap, SimpleMOSFET, PWLWave, PulseWave, SinWave, Wave)
from .circuit import (MNASpec, MNACircuit, Params, alter, with_mode, with_temp, with_gshunt,
                      with_srcfact)
from .sweeps import (Sweep, ProductSweep, TandemSweep, SerialSweep, CircuitSweep, SweepResult,
                     sweepvars, split_axes, sweepify, find_param_ranges)
from .lowering import lower, lower_circuit, LoweredCircuit, StructuralSweepError
from .analysis import (dc, tran, DCSolution, TranSolution, CompiledSweep, compile_sweep,
                       expand_breakpoints, breakpoints, CedarTranOp, CedarUICOp, state_abstol)
from .verilog_a import va, VAModel, VAError
from .behavioral import BehavioralVoltageSource, BehavioralCurrentSource
from . import backend

__all__ = [n for n in dir() if not n.startswith("_")]
