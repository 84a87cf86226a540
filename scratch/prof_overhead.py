import sys, time, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cadnip_b200 as cb
from cadnip_b200.workloads import clipper_sweep
cs = clipper_sweep(256, 256)
params, P = cs.lane_params()
lc = cb.lower(cs.builder, params, cb.MNASpec(mode="tran"), P=P)
comp = cb.CompiledSweep(lc, cb.MNASpec(mode="tran"))
t0=time.perf_counter(); comp.specialize(1e-6, "be"); print("specialize", time.perf_counter()-t0)
save=[lc.index_of("out")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def loop(tag, n=6, do_flush=False):
    for i in range(n):
        torch.cuda.synchronize()
        t0=time.perf_counter()
        wave = comp.tran((0,2e-3), 1e-6, method="be", save_idxs=save, save_every=10)
        t1=time.perf_counter()
        st = comp.handle.stats()
        wave.free()
        t2=time.perf_counter()
        if do_flush:
            flush.zero_(); torch.cuda.synchronize()
        t3=time.perf_counter()
        print(f"{tag}: tran wall {1e3*(t1-t0):.2f} ms  kernel {st['tran_kernel_ms']:.2f} dc {st['dc_kernel_ms']:.2f}  free {1e3*(t2-t1):.2f} flush {1e3*(t3-t2):.2f} ms launches {st['launches']}")
loop("plain")
loop("flush", do_flush=True)
p = subprocess.Popen(["nvidia-smi","--query-gpu=clocks.sm","--format=csv,noheader","-lms","100"], stdout=subprocess.DEVNULL)
time.sleep(1.0)
loop("smi")
p.terminate()
