#!/usr/bin/env python
"""bench.py -- N-pair fwd+bwd samples/sec at B=8192, D=512 on 1/2/4/8 B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16x2|bf16x3|bf16]

One "step" = Forward_gpu + Backward_gpu of the NPairMultiClassLoss layer over one batch of synthetic L2-normalised
embeddings (B/2 classes x 2 images, usage-block mining of usage/def.prototxt:137-146), the global batch B=8192
sharded by anchor over the N ranks (strong scaling: Q = B/N rows per rank, all N columns).

  value    : whole-job samples/sec with the inputs resident in HBM, through the C ABI (npair_forward + npair_backward,
             includes the all-gather / reduce-scatter and the five host scalars), CUDA-event timed, max over ranks.
  e2e      : same metric through the reference-facing plugin surface (the Caffe-style layer in npairloss_b200/caffe_shim)
             with HOST bottoms: H2D of features+labels from pinned memory and D2H of the gradient and tops inside
             the timed region.
  roofline : dominant kernel (similarity GEMM with fused statistics), live CUDA-event duration from a profiled pass.
  cpu_baseline : the oracle (a "port": the reference has no CPU path and cannot be compiled here) on a bounded sample.

--impl reference times the CPU oracle (faithful sorts, all host threads) on the same config; rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "N-pair fwd+bwd samples/sec at B=8192,D=512"
PRECS = {"bf16x3": 0, "bf16": 1, "fp16x2": 2}
MMA_PASSES = {"bf16x3": 6, "bf16": 1, "fp16x2": 3}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    sm_max_mhz=d.get("sm_max_mhz"), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def host_threads():
    """Threads for the CPU oracle: one per physical core (the sort-heavy oracle loses 4x with SMT oversubscription:
    683 vs 2690 samples/s on the 128-thread gpurun box).  torchrun exports OMP_NUM_THREADS=1, so never rely on the default."""
    n = os.cpu_count() or 1
    return max(1, n // 2) if n >= 32 else n


def run_reference(args, B, D, mining, noise):
    """CPU baseline arm: the oracle with the reference's unconditional sorts on all host threads.  Each step is the
    rank-0 block of the 8-way anchor sharding (1024 anchors x 8192 database), a bounded sample of the same workload."""
    from npairloss_b200 import synth
    from oracle import oracle_lib as o
    o.build()
    x, lab = synth.make_inputs(B, D, 20171225 + 5, noise=noise)
    world_s = 8
    Qs = B // world_s
    cores = host_threads()
    cfg = o.make_config(Qs, D, world=world_s, rank=0, accum_double=0, faithful_sorts=1, num_threads=cores, **mining)
    L = o.lib()
    buf = np.zeros(L.npo_state_floats(C.byref(cfg)), dtype=np.float32)
    st = o.NpoState()
    L.npo_state_bind(C.byref(cfg), o._fp(buf), C.byref(st))
    tops = np.zeros(5, np.float32)
    ld = np.zeros((Qs, D), np.float32)
    td = np.zeros((B, D), np.float32)

    def step():
        e = L.npo_forward(C.byref(cfg), o._fp(x), o._fp(lab), None, C.byref(st), o._fp(tops))
        e2 = L.npo_backward_partial(C.byref(cfg), o._fp(x), C.byref(st), C.c_float(1.0), o._fp(ld), o._fp(td))
        assert e == 0 and e2 == 0

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = Qs / dt
    sample = f"rank-0 block of the 8-way anchor sharding: {Qs} anchors x {B} database x D={D}, faithful sorts, fp32 accumulate"
    out = {"metric": METRIC, "impl": "reference", "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"HL: B={B}, D={D}, {B // 2} classes x 2, {args.mining_desc}, loss_weight 1",
                      "global_batch": B, "feature_dim": D, "noise": noise},
           "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The contract is ONE JSON line on stdout; libraries (NCCL's version banner, torchrun) also write to fd 1, so the
    process's stdout is pointed at stderr for the whole run and the result goes to the saved descriptor."""
    line = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16x2", choices=list(PRECS))
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # the headline is the reference's own usage block; "rand" (RAND/RAND, cheapest: no selects) and "relative"
    # (LOCAL RELATIVE_HARD both sides, diffsn -0.3: a radix select per row and side, costliest) are the two other settings it
    # asks to be reported.  Only "usage" is the BASELINE.json metric.
    ap.add_argument("--mining", default="usage", choices=["usage", "rand", "relative"])
    # extra, opt-in measurement (adds the key "e2e_prefetch"; "e2e" is unchanged): double-buffered host bottoms whose H2D copy
    # for step k+1 runs on a copy stream while step k computes -- what Caffe's prefetching data layers do
    ap.add_argument("--e2e-prefetch", action="store_true")
    # opt-in: the device-resident step calls npair_forward_backward (one host synchronisation per step) instead of
    # npair_forward + npair_backward
    ap.add_argument("--fused-step", action="store_true")
    args = ap.parse_args()

    from npairloss_b200 import synth
    B, D = args.batch, args.dim
    noise = synth.CONFIGS["HL"]["noise"]
    if args.mining == "usage":
        mining = dict(synth.USAGE_MINING)
        mining_desc = "usage-block mining (AP GLOBAL RELATIVE_HARD identsn -0.0, AN LOCAL HARD margin_diff -0.05)"
    elif args.mining == "rand":
        mining = dict(synth.DEFAULT_MINING)
        mining_desc = "RAND/RAND mining (proto defaults: every pair selected)"
    else:
        mining = dict(synth.USAGE_MINING, ap_region=synth.LOCAL, ap_method=synth.RELATIVE_HARD, an_region=synth.LOCAL,
                      an_method=synth.RELATIVE_HARD, identsn=-0.3, diffsn=-0.3, margin_diff=0.0)
        mining_desc = "LOCAL RELATIVE_HARD / RELATIVE_HARD mining, identsn = diffsn = -0.3 (per-row radix selects on both sides)"
    args.mining_desc = mining_desc
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        # every step is the bounded rank-0 sample (~0.5 s on the 128-thread gpurun host): K + W = 60 steps end within a minute
        if rank == 0:
            run_reference(args, B, D, mining, noise)
        return

    import torch
    import torch.distributed as dist
    from npairloss_b200 import capi

    assert torch.cuda.is_available(), "bench.py needs a B200 (the product path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert B % world == 0
    Q = B // world
    peaks = load_peaks()

    # identical bytes on every rank; each rank keeps its own row block
    x, lab = synth.make_inputs(B, D, 20171225 + 5, noise=noise)
    xl = np.ascontiguousarray(x[rank * Q:(rank + 1) * Q])
    ll = np.ascontiguousarray(lab[rank * Q:(rank + 1) * Q])

    nccl_id = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    cfg = capi.make_config(Q, D, world=world, rank=rank, sim_precision=PRECS[args.precision], device=local_rank, **mining)
    ctx = capi.Context(cfg, nccl_id)

    stream = torch.cuda.current_stream()
    d_x = torch.from_numpy(xl).to(dev)
    d_l = torch.from_numpy(ll).to(dev)
    d_g = torch.empty_like(d_x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        if args.fused_step:
            return ctx.forward_backward(d_x, d_l, 1.0, d_g)
        tops = ctx.forward(d_x, d_l)
        ctx.backward(1.0, d_g)
        return tops

    # ---------------- device-resident timed region ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # nvidia-smi takes ~0.1 s to produce its first line: start it ahead of the warm-up
    for _ in range(args.warmup):
        tops = step_device()
    barrier()
    n_before = len(sampler.rows)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # L2 rule: the step's working set is dominated by the Q x N fp32 similarity block.  When it is larger than twice the 126 MB
    # L2 nothing survives from one step to the next (N = 1: 268 MB); otherwise (sharded runs) a 252 MB device buffer is rewritten
    # before every step and the steps are timed one by one with their own event pair (the flush is outside the pairs).
    L2_BYTES = 126 << 20
    need_flush = 4 * Q * B < 2 * L2_BYTES
    flush_buf = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev) if need_flush else None
    barrier()
    if not need_flush:
        e0.record(stream)
        launches0 = capi.kernel_launches()
        for _ in range(args.steps):
            tops = step_device()
        launches_timed = capi.kernel_launches() - launches0
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
    else:
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        launches0 = capi.kernel_launches()
        for ea, eb in pairs:
            flush_buf.fill_(1)                 # same stream as the step: ordered before it, evicts the whole L2
            ea.record(stream)
            tops = step_device()
            eb.record(stream)
        launches_timed = capi.kernel_launches() - launches0
        barrier()
        ms_total = float(sum(ea.elapsed_time(eb) for ea, eb in pairs))
    clocks = None
    if rank == 0:
        in_region = len(sampler.rows) - n_before
        # a 50-step timed region lasts ~30 ms, shorter than nvidia-smi's sampling period: keep the SAME load running
        # (untimed) until a handful of samples exist, and say how many fell inside the timed region itself
        t_top = time.perf_counter()
        while len(sampler.rows) - n_before < 6 and time.perf_counter() - t_top < 2.0 and world == 1:
            step_device()
        clocks = sampler.stop()
        clocks["samples_in_timed_region"] = in_region
        clocks["note"] = "sampled with nvidia-smi -lms 25 from warm-up through the timed region and an untimed continuation of the same load"
    ms_t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = ms_t.item() / args.steps
    value = B / (ms_step * 1e-3)

    # ---------------- end-to-end through the plugin surface with host buffers ----------------
    # The Caffe-style layer (npairloss_b200/caffe_shim) on HOST blobs: every step a data layer hands a new batch through
    # mutable_cpu_data() (-> H2D of features + labels from pinned memory inside Blob::gpu_data()), Forward, Backward, and the
    # solver reads bottom diff on the host (-> D2H inside Blob::cpu_diff()); the five tops land on the host as well.
    from npairloss_b200 import caffe_layer
    nccl_id2 = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id2 = bytes(idt.cpu().numpy().tobytes())
    layer = caffe_layer.Layer(caffe_layer.layer_prototxt(mining, 5), Q, D, 1, 1, world=world, rank=rank, nccl_id=nccl_id2,
                              sim_precision=PRECS[args.precision])
    layer.bottom_data(0)[:] = xl.ravel()
    layer.bottom_data(1)[:] = ll
    e2e_api = ("caffe_shim NPairMultiClassLossLayer: Layer::Forward + Layer::Backward on host bottoms (prototxt-configured); tops read on "
               "the host every step, gradient left in bottom[0]'s device diff (fetched once after the loop for the consistency check)")

    # e2e step: H2D of the batch (pinned host blobs -> device, inside Forward), Forward with its host read of the five tops,
    # Backward.  The gradient stays in bottom[0]'s DEVICE diff, where the upstream layer consumes it in a net (the contract's
    # per-step device->host read is the step's result: loss + retrieval tops); it is fetched once after the loop for the check.
    for _ in range(max(3, args.warmup // 2)):
        tops_e = layer.step_host(read_gradient=False)
    barrier()
    e0.record(stream)
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        tops_e = layer.step_host(read_gradient=False)
    e1.record(stream)
    barrier()
    t_host = time.perf_counter() - t_host0
    # the shim's Blob copies are synchronous on the legacy stream; CUDA events on torch's stream still bracket them because
    # every step ends with a blocking D2H.  Use the larger of event time and host wall time.
    ms_e = torch.tensor([max(e0.elapsed_time(e1), t_host * 1e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e_value = B / (ms_e.item() / args.steps * 1e-3)
    h2d = Q * D * 4 + Q * 4
    d2h = 5 * 4 + 4                       # five tops + the error word, read from mapped pinned memory by Forward
    grad_e2e = layer.bottom_diff().copy()
    e2e_prefetch = None
    if args.e2e_prefetch:
        layer.prefetch_enable()
        for sset in (0, 1):
            layer.set_data(sset, 0)[:] = xl.ravel()
            layer.set_data(sset, 1)[:] = ll
        layer.prefetch(0)
        k = 0
        for _ in range(max(4, args.warmup // 2)):
            layer.prefetch((k + 1) & 1); layer.step_set(k & 1); k += 1
        barrier()
        t_p0 = time.perf_counter()
        for _ in range(args.steps):                 # one H2D (of the NEXT step's batch) is issued inside every timed step
            layer.prefetch((k + 1) & 1); tops_p = layer.step_set(k & 1); k += 1
        barrier()
        t_p = torch.tensor([time.perf_counter() - t_p0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_p, op=dist.ReduceOp.MAX)
        grad_p = layer.set_diff((k - 1) & 1).copy()
        e2e_prefetch = {"value": B / (t_p.item() / args.steps), "unit": "samples/s", "h2d_bytes_per_step": Q * D * 4 + Q * 4,
                        "d2h_bytes_per_step": 5 * 4 + 4, "timing": "host wall clock around the loop, device synchronised on both sides",
                        "api": "same layer calls on two bottom sets; SyncedMemory::async_gpu_push of the next batch on a copy stream",
                        "gradient_matches_device_path": bool(np.linalg.norm(grad_p - grad_e2e) <= 1e-6 * max(np.linalg.norm(grad_e2e), 1e-30)),
                        "loss": tops_p[0]}
    layer.close()
    torch.cuda.synchronize()
    grad_dev = d_g.cpu().numpy()
    e2e_consistent = bool(np.linalg.norm(grad_e2e - grad_dev) <= 1e-6 * max(np.linalg.norm(grad_dev), 1e-30))

    # ---------------- profiled pass: per-phase CUDA events (roofline of the dominant kernel) ----------------
    ctx.profile_enable(True)
    phases = np.zeros(9)
    nprof = min(args.steps, 10)
    for _ in range(nprof):
        step_device()
        phases += np.array(ctx.profile_read())
    phases /= nprof
    ctx.profile_enable(False)
    N = B
    sim_ms, grad_ms = float(phases[2]), float(phases[6])
    flops_alg = 2.0 * Q * N * D                                   # one Q x N x D contraction per launch, both GEMMs
    peak_tf = peaks["tf_sustained"]
    passes = MMA_PASSES[args.precision]
    # tiles the similarity kernel really issues: world == 1 computes only the 128 x 256 tiles that touch the upper triangle
    tm, tn = (Q + 127) // 128, (N + 255) // 256
    sym_frac = (sum(tn - mb // 2 for mb in range(tm)) / float(tm * tn)) if world == 1 else 1.0

    def tensor_roofline(kernel, ms, issued_factor, extra):
        ach = flops_alg / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        return {"kernel": kernel, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "mma_passes": passes, "issued_tflops": ach * issued_factor, "frac_of_issued_mma": ach * issued_factor / peak_tf,
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "duration_ms": ms, "traffic": None,
                "note": "achieved = algorithmic 2*Q*N*D flops / live CUDA-event duration of the phase; the fp32-faithful modes issue "
                        "mma_passes bf16-rate MMA passes per algorithmic flop; issued_tflops / frac_of_issued_mma count the MMA work "
                        "the kernel really issues (" + extra + ")"}

    r_sim = tensor_roofline("split_gemm_kernel<EPI_SIM*> CTA-pair tcgen05 similarity GEMM + fused row statistics", sim_ms, passes * sym_frac,
                            f"passes x {sym_frac:.3f} of the tiles: symmetric tile list" if world == 1 else "passes x all tiles")
    r_grad = tensor_roofline("fused_grad_kernel CTA-pair tcgen05 gradient GEMM, weights produced into tensor memory", grad_ms, passes,
                             "passes x all tiles; includes the split-K reduce when Q = B/world leaves few tiles")
    # DRAM bytes per launch from the committed ncu --set full capture (profiles/r01_traffic.json), when it is this workload
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_traffic.json")) as f:
            tr = json.load(f)
        w = tr["workload"]
        if (w["B"], w["D"], w["precision"], w["world"]) == (B, D, args.precision, world):
            r_sim["traffic"] = tr["kernels"]["sim_gemm"]["dram_bytes"]
            r_grad["traffic"] = tr["kernels"]["grad_gemm"]["dram_bytes"]
            r_sim["traffic_source"] = r_grad["traffic_source"] = "profiles/r01_traffic.json (ncu --set full, bytes per launch)"
    except Exception:
        pass
    roofline = r_grad if grad_ms >= sim_ms else r_sim              # the dominant kernel of the step
    roofline_other = [r_sim if grad_ms >= sim_ms else r_grad]
    phase_names = ["fwd_allgather", "operand_prep", "sim_gemm", "thresholds_select", "row_pass_finalize", "weight_build", "grad_gemm",
                   "grad_gemm_T", "bwd_exchange"]
    phase_ms = {n: float(v) for n, v in zip(phase_names, phases)}
    # memory-bound kernels: algorithmic bytes = one fp32 pass over the Q x N block
    sbytes = 4.0 * Q * N
    hbm = {"row_pass_GBs": sbytes / (phases[4] * 1e-3) / 1e9 if phases[4] > 0 else None,
           "weight_build_GBs": sbytes / (phases[5] * 1e-3) / 1e9 if phases[5] > 0 else None,
           "hbm_peak_GBs": peaks["hbm_gbs"]}

    # kernels launched inside the timed region by OUR library: counted by the library itself (npair_kernel_launches)
    gpu_launches = launches_timed

    # ---------------- CPU baseline (rank 0, bounded sample) ----------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle_lib as o
        o.build()
        world_s = 8
        Qs = B // world_s
        ocfg = o.make_config(Qs, D, world=world_s, rank=0, accum_double=0, faithful_sorts=1, num_threads=host_threads(), **mining)
        L = o.lib()
        buf = np.zeros(L.npo_state_floats(C.byref(ocfg)), dtype=np.float32)
        st = o.NpoState()
        L.npo_state_bind(C.byref(ocfg), o._fp(buf), C.byref(st))
        t5 = np.zeros(5, np.float32)
        ld = np.zeros((Qs, D), np.float32)
        td = np.zeros((B, D), np.float32)
        reps, t0 = 0, time.perf_counter()
        while reps < 2 or (time.perf_counter() - t0 < 10 and reps < 8):
            assert L.npo_forward(C.byref(ocfg), o._fp(x), o._fp(lab), None, C.byref(st), o._fp(t5)) == 0
            assert L.npo_backward_partial(C.byref(ocfg), o._fp(x), C.byref(st), C.c_float(1.0), o._fp(ld), o._fp(td)) == 0
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": Qs / dt, "unit": "samples/s", "cores": host_threads(), "kind": "port",
               "sample": f"rank-0 block of the 8-way anchor sharding: {Qs} anchors x {B} database x D={D}, {reps} reps, "
                         "faithful unconditional sorts, fp32 accumulate, OpenMP, one thread per physical core"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": {"fp16x2": "f32 (3-pass fp16-split tcgen05, f32 accumulate)", "bf16x3": "f32 (6-pass bf16-split tcgen05, f32 accumulate)",
                         "bf16": "bf16 (f32 accumulate)"}[args.precision],
               "data": "synthetic",
               "config": {"workload": f"HL: B={B}, D={D}, {B // 2} classes x 2, {args.mining_desc}, loss_weight 1",
                          "global_batch": B, "feature_dim": D,
                          "rows_per_rank": Q, "sharding": f"anchor-sharded x{world}", "precision": args.precision, "noise": noise,
                          "step_call": "npair_forward_backward (one host sync)" if args.fused_step else "npair_forward + npair_backward",
                          "l2": (f"flushed: a {2 * L2_BYTES >> 20} MB device buffer is rewritten before every timed step (per-rank S = "
                                 f"{4 * Q * N / 1e6:.0f} MB fp32 would otherwise stay in the 126 MB L2); steps timed one by one, flush excluded"
                                 if need_flush else
                                 f"not flushed: per-step working set (S {4 * Q * N / 1e6:.0f} MB fp32 + operand pieces) is more than twice the 126 MB L2")},
               "clocks": clocks, "roofline": roofline, "roofline_other": roofline_other, "e2e_prefetch": e2e_prefetch, "phase_ms": phase_ms, "hbm_kernels": hbm,
               "cpu_baseline": cpu,
               "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": e2e_api,
                       "gradient_matches_device_path": e2e_consistent},
               "gpu_launches": gpu_launches,
               "tops": {"loss": tops[0], "top1": tops[1], "top5": tops[2], "top10": tops[3], "feature_asum": tops[4]}}
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
