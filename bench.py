#!/usr/bin/env python
"""bench.py -- headline benchmark of the Vivim Temporal-Mamba hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips C]

Metric (BASELINE.json): selective-scan fwd+bwd algorithmic GB/s at the Vivim stage-1 shape
(configs[1]: L = 5*64*64 = 20480 tokens, d_inner 128, d_state 16, bf16 I/O, fp32 state).
A "step" = one forward + one backward of the fused selective scan over `clips` clips per GPU.

  value     device-resident inputs, the 6 scan kernels (+ accumulator zero-fill and the dB/dC cast that
            the reference's C++ shim also performs) replayed from a CUDA graph, timed with CUDA
            events over exactly K steps; algorithmic bytes = (11T + 6S) per clip.
  e2e       the same op through the public API (mamba_ssm.ops.selective_scan_interface.
            selective_scan_fn + autograd backward) with HOST (pinned) inputs copied in and all
            results copied out inside the timed region.
  roofline  the dominant kernel (seg_bwd_kernel) timed alone with CUDA events.
  cpu_baseline / --impl reference   the torch port of selective_scan_ref (oracle/torch_port.py)
            timed on the host cores over a bounded sample of the same workload.

Multi-GPU (torchrun): clips are independent, every rank scans its own clips, no data-path
collective; value = total bytes of all ranks / max-over-ranks time ("weak" scaling).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stage-1 Temporal Mamba block of Vivim at image 256, clip_length 5
D_INNER, SEQLEN, D_STATE = 128, 5 * 64 * 64, 16
L2_BYTES = 126 * 1024 * 1024
METRIC = "mamba_scan_fwd_bwd_GBps_stage1"


def algo_bytes(clips, seqlen=SEQLEN, elem=2):
    """(fwd, bwd) algorithmic bytes: fwd 4T+2S, bwd 7T+4S."""
    T = clips * D_INNER * seqlen * elem
    S = clips * D_STATE * seqlen * elem
    return 4 * T + 2 * S, 7 * T + 4 * S


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first_sample(self, timeout=3.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """samples taken from now on count as 'under load'"""
        self.t_mark = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thr.join(timeout=1)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, r in self.rows:
            if ts < t_mark:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arms (reference / cpu_baseline): torch port of selective_scan_ref on a bounded sample
# ------------------------------------------------------------------------------------------------
CPU_SAMPLE_L = 512


def cpu_sample_inputs(seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    L = CPU_SAMPLE_L
    A = -torch.arange(1, D_STATE + 1, dtype=torch.float32).repeat(D_INNER, 1)
    mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16).float()  # noqa: E731
    t = dict(u=mk(1, D_INNER, L), delta=0.5 * mk(1, D_INNER, L), A=A, B=mk(1, D_STATE, L), C=mk(1, D_STATE, L),
             D=torch.ones(D_INNER), z=mk(1, D_INNER, L), bias=torch.rand(D_INNER, generator=g) - 4.0,
             dout=mk(1, D_INNER, L))
    return t


def cpu_step(t):
    """fwd + autograd bwd of the torch port -- the reference's own CPU path
    (selective_scan_ref + torch autograd, mamba/tests/ops/test_selective_scan.py:97-124)."""
    from oracle.torch_port import selective_scan_port
    leaves = {k: t[k].clone().requires_grad_() for k in ("u", "delta", "A", "B", "C", "D", "z", "bias")}
    out = selective_scan_port(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                              z=leaves["z"], delta_bias=leaves["bias"], delta_softplus=True)
    out.backward(t["dout"])
    return out


def time_cpu(steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    t = cpu_sample_inputs()
    for _ in range(warmup):
        cpu_step(t)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(t)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    fwd_b, bwd_b = algo_bytes(1, CPU_SAMPLE_L)
    return (fwd_b + bwd_b) / dt / 1e9, dt, torch.get_num_threads()


def cpu_baseline_obj(value, cores):
    return {"value": value, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": f"selective_scan_ref torch port fwd+autograd-bwd, 1 clip, first {CPU_SAMPLE_L} of {SEQLEN} "
                      f"tokens, d_inner {D_INNER}, d_state {D_STATE}, fp32 (the ref backward is O(L^2), "
                      f"so the full length does not finish)"}


def workload_config(clips, extra=None):
    cfg = {"workload": "single Temporal Mamba block selective-scan fwd+bwd, Vivim stage-1 shape (BASELINE configs[1])",
           "clips_per_gpu": clips, "seqlen": SEQLEN, "d_inner": D_INNER, "d_state": D_STATE, "io": "bf16", "state": "fp32"}
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args, rank, world):
    if rank != 0:
        return
    value, dt, cores = time_cpu(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.clips, {"note": "CPU arm; each step is the bounded sample in cpu_baseline.sample"}),
            "cpu_baseline": cpu_baseline_obj(value, cores),
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ScanSet:
    """One set of device-resident inputs + preallocated outputs/workspaces + prebuilt C-ABI args."""

    def __init__(self, clips, device, seed):
        import torch
        from vivim_b200 import _lib
        g = torch.Generator(device="cpu").manual_seed(seed)
        bf = torch.bfloat16
        B_, D_, L_, N_ = clips, D_INNER, SEQLEN, D_STATE
        U = (L_ + _lib.VV_SCAN_SEGMENT - 1) // _lib.VV_SCAN_SEGMENT
        r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
        self.host = dict(u=r(B_, D_, L_).to(bf), delta=(0.5 * r(B_, D_, L_)).to(bf), z=r(B_, D_, L_).to(bf),
                         B=r(B_, 1, N_, L_).to(bf), C=r(B_, 1, N_, L_).to(bf), dout=r(B_, D_, L_).to(bf))
        self.t = {k: v.to(device) for k, v in self.host.items()}
        # module init of mamba_simple.py:99-117: A = -(1..N), D = 1, dt bias = softplus^-1(U[1e-3, 1e-1])
        dt0 = torch.exp(torch.rand(D_, generator=g) * (torch.log(torch.tensor(0.1)) - torch.log(torch.tensor(1e-3)))
                        + torch.log(torch.tensor(1e-3))).clamp(min=1e-4)
        self.p = dict(A=-torch.arange(1, N_ + 1, dtype=torch.float32).repeat(D_, 1).to(device),
                      D=torch.ones(D_, device=device), bias=(dt0 + torch.log(-torch.expm1(-dt0))).to(device))
        e = lambda: torch.empty(B_, D_, L_, dtype=bf, device=device)  # noqa: E731
        self.out_z, self.du, self.ddelta, self.dz = e(), e(), e(), e()
        n_bc = B_ * N_ * L_
        self.acc = torch.zeros(2 * n_bc + D_ * N_ + 2 * D_, dtype=torch.float32, device=device)
        self.dBC16 = torch.empty(2, B_, 1, N_, L_, dtype=bf, device=device)
        self.agg = torch.empty(B_, D_, U, N_, 2, dtype=torch.float32, device=device)
        self.chk = torch.empty(B_, D_, U, N_, dtype=torch.float32, device=device)
        self.radj = torch.empty(B_, D_, U, N_, dtype=torch.float32, device=device)
        self.last = torch.empty(B_, D_, N_, dtype=torch.float32, device=device)
        a = _lib.ScanArgs()
        t, p = self.t, self.p
        a.u, a.delta, a.z, a.Bm, a.Cm, a.dout = (t[k].data_ptr() for k in ("u", "delta", "z", "B", "C", "dout"))
        a.A, a.D, a.delta_bias = p["A"].data_ptr(), p["D"].data_ptr(), p["bias"].data_ptr()
        a.out_z, a.du, a.ddelta, a.dz = (x.data_ptr() for x in (self.out_z, self.du, self.ddelta, self.dz))
        a.last_state, a.agg, a.chk, a.radj = (x.data_ptr() for x in (self.last, self.agg, self.chk, self.radj))
        base = self.acc.data_ptr()
        a.dB, a.dC = base, base + 4 * n_bc
        a.dB_io, a.dC_io = self.dBC16[0].data_ptr(), self.dBC16[1].data_ptr()
        a.dA = base + 8 * n_bc
        a.dD = a.dA + 4 * D_ * N_
        a.ddelta_bias = a.dD + 4 * D_
        a.batch, a.dim, a.seqlen, a.dstate, a.ngroups = B_, D_, L_, N_, 1
        for name in ("u", "delta", "z", "out", "outz", "dout", "du", "ddelta", "dz"):
            setattr(a, name + "_bs", D_ * L_)
            setattr(a, name + "_ds", L_)
        a.A_ds, a.A_ns = N_, 1
        a.B_bs = a.C_bs = N_ * L_
        a.B_gs = a.C_gs = N_ * L_
        a.B_ns = a.C_ns = L_
        a.io_dtype, a.delta_softplus, a.zero_accumulators = _lib.VV_BF16, 1, 1
        self.args = a
        self.n_bc = n_bc

    def input_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.t.values())


def launch_step(s, lib, stream):
    """scan fwd (3 kernels) -> scan bwd (3 kernels, the first one zero-fills the fp32 accumulators) -> dB/dC cast to bf16 (a fourth kernel of vv_scan_bwd, chained with programmatic dependent launch)."""
    from vivim_b200 import _lib
    _lib.check(lib.vv_scan_fwd(ctypes.byref(s.args), ctypes.c_void_p(stream)), "vv_scan_fwd")
    _lib.check(lib.vv_scan_bwd(ctypes.byref(s.args), ctypes.c_void_p(stream)), "vv_scan_bwd")
    return 7


def time_events(fn, iters, torch):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from vivim_b200 import _lib, build as vbuild
    from vivim_b200.sharding import aggregate_throughput, max_over_ranks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    vbuild.build()
    lib = _lib.lib()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    clips = args.clips
    fwd_b, bwd_b = algo_bytes(clips)
    # inputs larger than L2: rotate over enough input sets that a set is evicted before it is reused
    probe = ScanSet(clips, device, seed=1000 * rank)
    n_sets = max(2, -(-2 * L2_BYTES // probe.input_bytes()) + 1)
    sets = [probe] + [ScanSet(clips, device, seed=1000 * rank + i) for i in range(1, n_sets)]
    l2_policy = (f"rotating {n_sets} device-resident input sets ({n_sets * probe.input_bytes() / 2**20:.0f} MiB) "
                 f"> 126 MiB L2, so every step reads its inputs from HBM")

    # ---- capture one CUDA graph per input set (launch-bound inner loop -> graph replay)
    side = torch.cuda.Stream(device)
    graphs = []
    with torch.cuda.stream(side):
        for s in sets:                      # warm-up outside capture (sets func attributes, fills workspaces)
            launch_step(s, lib, side.cuda_stream)
    side.synchronize()
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            launch_step(s, lib, torch.cuda.current_stream().cuda_stream)
        graphs.append(g)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = lambda i: graphs[i % n_sets].replay()  # noqa: E731
    with ClockSampler(local_rank) as clk:
        clk.wait_first_sample()
        for i in range(args.warmup):
            step(i)
        barrier()
        clk.mark()
        elapsed = time_events(lambda i: step(i + args.warmup), args.steps, torch)
        barrier()
    elapsed = max_over_ranks(elapsed, device)
    value = aggregate_throughput((fwd_b + bwd_b) * args.steps, world, elapsed) / 1e9
    launches = 7 * args.steps

    # ---- the six kernels one by one (pass mask), rotating sets, CUDA events on the launch stream
    stream = torch.cuda.current_stream().cuda_stream
    passes = {}
    reps = max(10, min(50, args.steps))
    for bwd in (0, 1):
        fn = lib.vv_scan_bwd if bwd else lib.vv_scan_fwd
        for bit, name in ((1, "agg"), (2, "carry"), (4, "main")):
            for s in sets:
                s.args.pass_mask = bit
            run = lambda i: _lib.check(fn(ctypes.byref(sets[i % n_sets].args), ctypes.c_void_p(stream)), "scan pass")  # noqa: E731
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            passes[("bwd_" if bwd else "fwd_") + name] = time_events(run, reps, torch) / reps
    for s in sets:
        s.args.pass_mask = 0
    launches += 6 * (reps + 3)
    t_main = passes["bwd_main"]
    peak, peak_src = measured_peak()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("seg_bwd_kernel_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "seg_bwd_kernel", "achieved": bwd_b / t_main / 1e9, "peak": peak,
                "unit": "GB/s", "frac": bwd_b / t_main / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bwd_b, "us_per_launch": t_main * 1e6,
                "note": "N=16 states per (channel, token): ~26 issue slots, 1 MUFU and 6.6 shared-memory/shuffle wavefront "
                        "bytes per state-step make this kernel issue / LSU bound, not HBM bound, on B200 (DESIGN.md section 5)"}

    # ---- conv1d fwd / bwd at the same shape (x = first half of xz), for the record
    conv = bench_conv(clips, device, torch)
    launches += conv.pop("_launches")

    # ---- end to end: public API, host (pinned) buffers in, all results out, every step
    e2e_steps = max(4, min(args.steps, 40))
    e2e = bench_e2e(sets[0], e2e_steps, world, device, torch, dist)
    launches += 6 * (3 * e2e_steps + 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # Second roofline, the one that actually binds these kernels: MUFU.EX2 lane-operations per pass (one per state-step
    # for the decay, plus softplus / sigmoid in the per-position pre-pass) against the measured MUFU rate
    # (scripts/microbench/mufu_rate.cu: 15.57 lanes/clk/SM on B200) at the SM clock sampled during the run.
    ss = clips * D_INNER * SEQLEN * D_STATE            # state-steps
    pos = clips * D_INNER * SEQLEN                     # (channel, position) pairs
    mufu_ops = {"fwd_agg": ss + 2 * pos, "fwd_main": ss + 2 * pos + 2 * pos, "bwd_agg": ss + 4 * pos,
                "bwd_main": ss * 9 // 8 + 5 * pos}
    sm_mhz = (clk.summary().get("sm_mhz") or 1965.0)
    mufu_peak = 15.57 * 148 * sm_mhz * 1e6
    mufu = {k: {"mufu_lane_ops": v, "floor_us": v / mufu_peak * 1e6, "frac_of_mufu_peak": v / mufu_peak / passes[k]}
            for k, v in mufu_ops.items()}
    line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(clips, {"l2_policy": l2_policy, "launch": "CUDA graph replay, one graph per input set"}),
            "hbm_frac_of_measured_peak": value / world / peak,
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clk.summary(),
            "kernel_us": {k: v * 1e6 for k, v in passes.items()},
            "mufu_roofline": {"peak_lane_ops_per_s": mufu_peak, "source": "scripts/microbench/mufu_rate.cu (15.57 lanes/clk/SM)",
                              "kernels": mufu},
            "fwd_GBps_kernels_only": fwd_b / (passes["fwd_agg"] + passes["fwd_carry"] + passes["fwd_main"]) / 1e9,
            "bwd_GBps_kernels_only": bwd_b / (passes["bwd_agg"] + passes["bwd_carry"] + passes["bwd_main"]) / 1e9,
            "conv1d": conv}
    if world == 1:
        cpu_v, _, cores = time_cpu(steps=2, warmup=1)
        line["cpu_baseline"] = cpu_baseline_obj(cpu_v, cores)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_conv(clips, device, torch):
    from vivim_b200 import causal_conv1d_cuda as ccc
    bf = torch.bfloat16
    T = clips * D_INNER * SEQLEN * 2
    n_sets = max(2, -(-2 * L2_BYTES // (3 * T)) + 1)
    xz = [torch.randn(clips, 2 * D_INNER, SEQLEN, device=device, dtype=bf) for _ in range(n_sets)]
    dout = [torch.randn(clips, D_INNER, SEQLEN, device=device, dtype=bf) for _ in range(n_sets)]
    w = torch.randn(D_INNER, 4, device=device)
    b = torch.randn(D_INNER, device=device)
    dxz = torch.empty_like(xz[0])
    reps = 30

    def fwd(i):
        ccc.causal_conv1d_fwd(xz[i % n_sets][:, :D_INNER], w, b, True)

    def bwd(i):
        ccc.causal_conv1d_bwd(xz[i % n_sets][:, :D_INNER], w, b, dout[i % n_sets], dxz[:, :D_INNER], True)

    res = {}
    for name, fn, nbytes in (("fwd", fwd, 2 * T), ("bwd", bwd, 3 * T)):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device)
        with torch.cuda.stream(side):
            for i in range(n_sets):
                fn(i)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            for i in range(n_sets):
                fn(i)
        torch.cuda.synchronize()
        g.replay()
        t = time_events(lambda i: g.replay(), reps, torch) / (reps * n_sets)
        res[name + "_GBps"] = nbytes / t / 1e9
        res[name + "_us"] = t * 1e6
    res["note"] = f"x = first half of xz ({clips},{2 * D_INNER},{SEQLEN}) bf16, K=4, SiLU; algorithmic bytes 2T fwd / 3T bwd"
    res["_launches"] = 2 * n_sets * (reps + 3)
    return res


def bench_e2e(s, steps, world, device, torch, dist, graphed=True):
    """selective_scan_fn + backward through the public API: every step copies its inputs in from pinned host
    memory and every result (out_z and all eight gradients) back out to pinned host memory.

    Host side kept lean, because at 0.12 ms of kernels per step the Python / launch overhead is what a user
    would otherwise measure: the inputs of a step are one flat pinned buffer (one H2D copy, the tensors are views
    of the device copy), forward + backward are replayed as CUDA graphs (vivim_b200.graphed.graph_module around
    selective_scan_fn -- the launches of this library are capture-safe), the nine results are gathered into one
    flat device buffer (one D2H copy).  Copies of neighbouring steps overlap the kernels: three streams, two
    buffer sets."""
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
    from vivim_b200.graphed import graph_module
    from vivim_b200.sharding import aggregate_throughput, max_over_ranks
    order = ("u", "delta", "A", "B", "C", "D", "z", "bias")
    src = dict(s.host)
    src.update({k: v.cpu() for k, v in s.p.items()})
    # flat layout: every tensor at a 256-byte aligned offset
    offs, total = {}, 0
    for k in order + ("dout",):
        offs[k] = total
        total += -(-src[k].numel() * src[k].element_size() // 256) * 256
    pin_in = torch.empty(total, dtype=torch.uint8, pin_memory=True)

    def views(flat):
        return {k: flat[offs[k]:offs[k] + src[k].numel() * src[k].element_size()].view(src[k].dtype).view(src[k].shape)
                for k in offs}

    for k, v in views(pin_in).items():
        v.copy_(src[k])
    h2d = sum(v.numel() * v.element_size() for v in src.values())
    depth = 2
    dev_flat = [torch.empty(total, dtype=torch.uint8, device=device) for _ in range(depth)]
    dev_in = [views(f) for f in dev_flat]

    def scan(u, delta, A, B, C, D, z, bias):
        return selective_scan_fn(u, delta, A, B, C, D, z=z, delta_bias=bias, delta_softplus=True)

    fn = scan
    if graphed:
        sample = tuple(dev_in[0][n].clone().requires_grad_() for n in order)
        fn = graph_module(scan, sample)
    s_in, s_cmp, s_out = (torch.cuda.Stream(device) for _ in range(3))
    ev_in = [torch.cuda.Event() for _ in range(depth)]
    ev_cmp = [torch.cuda.Event() for _ in range(depth)]
    ev_out = [torch.cuda.Event() for _ in range(depth)]
    out_bytes = h2d - src["dout"].numel() * src["dout"].element_size() + src["u"].numel() * src["u"].element_size()
    dev_out = [torch.empty(out_bytes, dtype=torch.uint8, device=device) for _ in range(depth)]
    host_out = [torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(depth)]

    def one(i):
        k = i % depth
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp[k])              # the kernels that last read this input set are done
            dev_flat[k].copy_(pin_in, non_blocking=True)
            ev_in[k].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[k])
            s_cmp.wait_event(ev_out[k])             # the previous results of this slot have left the device
            leaves = [dev_in[k][n].detach().requires_grad_() for n in order]
            out = fn(*leaves)
            out.backward(dev_in[k]["dout"])
            results = [out.detach()] + [x.grad for x in leaves]
            torch.cat([r.reshape(-1).view(torch.uint8) for r in results], out=dev_out[k])
            ev_cmp[k].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[k])
            host_out[k].copy_(dev_out[k], non_blocking=True)
            ev_out[k].record(s_out)

    for i in range(2 * depth):
        one(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    def block():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()                                 # current stream; the three streams start after it
        for st in (s_in, s_cmp, s_out):
            st.wait_event(e0)
        for i in range(steps):
            one(i)
        cur = torch.cuda.current_stream()
        for st in (s_in, s_cmp, s_out):
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3

    # three timed blocks of `steps` steps; the median block is reported (a single block is at the mercy of one
    # scheduling hiccup on the box)
    blocks = sorted(block() for _ in range(3))
    elapsed = blocks[1]
    elapsed = max_over_ranks(elapsed, device)
    fwd_b, bwd_b = algo_bytes(s.t["u"].shape[0])
    return {"value": aggregate_throughput((fwd_b + bwd_b) * steps, world, elapsed) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": out_bytes, "steps": steps, "ms_per_step": elapsed / steps * 1e3,
            "ms_per_step_blocks": [b / steps * 1e3 for b in blocks],
            "api": "mamba_ssm.ops.selective_scan_interface.selective_scan_fn + autograd backward"
                   + (", replayed as CUDA graphs (vivim_b200.graphed.graph_module)" if graphed else ""),
            "pipeline": "one flat H2D copy / kernels / one flat D2H copy on three streams, two buffer sets: copies of "
                        "step i+1 and i-1 overlap the kernels of step i"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--clips", type=int, default=1, help="clips per GPU (BASELINE configs[1]: 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.steps = 5 if args.steps is None else min(args.steps, 20)
        args.warmup = 1 if args.warmup is None else min(args.warmup, 3)
        run_reference(args, rank, world)
        return
    args.steps = 3000 if args.steps is None else args.steps
    args.warmup = 100 if args.warmup is None else max(args.warmup, 3)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
