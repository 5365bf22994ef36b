"""CPU baseline table of SURVEY.md 8(d): the faithful C++ oracle (unconditional sorts, fp32 accumulate) timed on THIS host for
the BASELINE.json shapes, single-threaded (the reference's host part is single-threaded) and on all cores.
    python oracle/time_oracle.py [--quick]        -> markdown rows on stdout
Test infrastructure: not imported by the product."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from npairloss_b200 import synth
from oracle import oracle_lib as o

def time_case(name, B, D, world, mining, noise, idx, threads, budget_s):
    x, lab = synth.make_inputs(B, D, 20171225 + idx, noise=noise)
    Q = B // world
    cfg = o.make_config(Q, D, world=world, rank=0, accum_double=0, faithful_sorts=1, num_threads=threads, **mining)
    L = o.lib()
    buf = np.zeros(L.npo_state_floats(C.byref(cfg)), dtype=np.float32)
    st = o.NpoState(); L.npo_state_bind(C.byref(cfg), o._fp(buf), C.byref(st))
    tops = np.zeros(5, np.float32); ld = np.zeros((Q, D), np.float32); td = np.zeros((B, D), np.float32)
    reps, t0 = 0, time.perf_counter()
    while reps < 1 or (time.perf_counter() - t0 < budget_s and reps < 5):
        assert L.npo_forward(C.byref(cfg), o._fp(x), o._fp(lab), None, C.byref(st), o._fp(tops)) == 0
        assert L.npo_backward_partial(C.byref(cfg), o._fp(x), C.byref(st), C.c_float(1.0), o._fp(ld), o._fp(td)) == 0
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return dt, Q

if __name__ == "__main__":
    o.build()
    quick = "--quick" in sys.argv
    ncore = os.cpu_count() or 1
    cases = [("C1", synth.CONFIGS["C1"], 1), ("C2", synth.CONFIGS["C2"], 1), ("C3", synth.CONFIGS["C3"], 1),
             ("C4 rank-0 block (k=8)", synth.CONFIGS["C4"], 8), ("HL rank-0 block (k=8)", synth.CONFIGS["HL"], 8), ("HL (k=1)", synth.CONFIGS["HL"], 1)]
    print(f"| config | rows x database x D | 1 thread: s/iter (samples/s) | {ncore} threads: s/iter (samples/s) |")
    print("|---|---|---|---|")
    for name, c, world in cases:
        if quick and c["B"] > 4096 and world == 1: continue
        cells = []
        for th in (1, ncore):
            dt, Q = time_case(name, c["B"], c["D"], world, c["mining"], c["noise"], c["idx"], th, 5.0 if not quick else 1.0)
            cells.append(f"{dt:.3g} ({Q / dt:.3g})")
        print(f"| {name} | {c['B'] // world} x {c['B']} x {c['D']} | {cells[0]} | {cells[1]} |", flush=True)
