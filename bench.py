#!/usr/bin/env python
"""bench.py -- headline benchmark of the Vivim Temporal-Mamba hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips C] [--dirs 1|3] [--no-vivim]

Metric (BASELINE.json): "Mamba scan fwd+bwd GB/s vs HBM peak; Vivim clips/sec at 1/2/4/8 B200".

  metric / value   selective-scan fwd+bwd algorithmic GB/s at the Vivim stage-1 shape (configs[1]: L = 5*64*64 = 20480
            tokens, d_inner 128, d_state 16, bf16 I/O, fp32 state).  A "step" = forward + backward of one scan over
            `clips` clips per GPU; algorithmic bytes = SURVEY.md 8(d): fwd 4T+2S, bwd 7T+4S.  Device-resident inputs,
            CUDA-graph replay, CUDA events over exactly K steps.
  block_three_directions   a Temporal Mamba block scans its tokens in three directions (mamba_simple.py:217-260); here
            they are ONE launch chain (direction blocks with per-block traversal order, B / C read from x_dbl rows, z /
            dout shared) -- the launch vivim_b200.mamba_block issues.  Bytes: 3 x the per-scan figure;
            `bytes_compulsory` is what the fused launch itself has to move.
  e2e       the headline workload through the reference-facing plugin boundary (vivim_b200.selective_scan_cuda.fwd / .bwd, the
            module that stands where the reference's pybind `selective_scan_cuda` stands; C ABI underneath) with HOST
            buffers: every step copies its inputs in from pinned host memory and all results back out.
  roofline  the dominant kernel (seg_bwd_kernel) timed alone with CUDA events.
  vivim_train_clips_per_s / vivim_infer_clips_per_s   the second half of the metric (configs[2] / configs[3]): the whole
            network, synthetic clips, random init, bf16 autocast, recall_focused_loss + AdamW, gradient all-reduce over
            NCCL when N > 1; one process per GPU, max-over-ranks time.
  ref_cuda_us   the UNMODIFIED reference kernels (baseline/_ref, built by baseline/build_ref.py) on the same B200.
  cpu_baseline / --impl reference   the CPU restatement of the reference path (oracle/oracle.c, O(L) forward and
            analytic backward, all host threads) on the SAME workload at the full length; the torch port of
            selective_scan_ref (the reference's own op-for-op CPU code, O(L^2) backward) is reported beside it at the
            sizes it finishes (forward at full L, fwd+bwd at L = 4096).

Multi-GPU (torchrun): clips are independent, every rank scans its own clips, no data-path collective; value = total
bytes of all ranks / max-over-ranks time ("weak" scaling).
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
D_INNER, SEQLEN, D_STATE, DT_RANK, NFRAMES = 128, 5 * 64 * 64, 16, 4, 5
DIRS = ("fwd", "rev", "frames")            # Mamba.forward v3 (mamba_simple.py:217-260)
L2_BYTES = 126 * 1024 * 1024
METRIC = "mamba_scan_fwd_bwd_GBps_stage1"


def algo_bytes(scans, seqlen=SEQLEN, elem=2):
    """(fwd, bwd) algorithmic bytes of `scans` direction scans: fwd 4T+2S, bwd 7T+4S each."""
    T = scans * D_INNER * seqlen * elem
    S = scans * D_STATE * seqlen * elem
    return 4 * T + 2 * S, 7 * T + 4 * S


def compulsory_bytes(clips, ndirs, seqlen=SEQLEN, elem=2):
    """What one fused launch has to move: z and dout are shared by the directions (read once)."""
    T = clips * D_INNER * seqlen * elem
    S = clips * D_STATE * seqlen * elem
    fwd = (2 * ndirs + 1 + ndirs) * T + 2 * ndirs * S            # u, delta per direction; z once; out_z per direction
    bwd = (2 * ndirs + 2 + 3 * ndirs) * T + 4 * ndirs * S        # + dout once; du, ddelta, dz per direction; dB, dC
    return fwd, bwd


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
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}



def workload_config(clips, ndirs):
    T_in = clips * D_INNER * SEQLEN * 2
    set_bytes = (2 * ndirs + 2) * T_in + clips * ndirs * SEQLEN * (DT_RANK + 2 * D_STATE) * 2
    n_sets = max(2, -(-2 * L2_BYTES // set_bytes) + 1)
    return {"workload": ("single Temporal Mamba block selective-scan fwd+bwd, Vivim stage-1 shape (BASELINE configs[1])" if ndirs == 1 else
                         "one Temporal Mamba block: selective-scan fwd+bwd of its %d direction scans in one launch chain, Vivim "
                         "stage-1 shape" % ndirs),
            "clips_per_gpu": clips, "directions": list(DIRS[:ndirs]), "seqlen": SEQLEN, "d_inner": D_INNER, "d_state": D_STATE,
            "io": "bf16", "state": "fp32",
            "l2_policy": f"rotating {n_sets} device-resident input sets ({n_sets * set_bytes / 2**20:.0f} MiB) > 126 MiB L2, "
                         f"so every step reads its inputs from HBM",
            "launch": "CUDA graph replay, one graph per input set"}, n_sets


# ------------------------------------------------------------------------------------------------
# CPU arms (reference / cpu_baseline)
# ------------------------------------------------------------------------------------------------
def block_inputs_numpy(clips, ndirs, seed=0):
    import numpy as np
    g = np.random.default_rng(seed)
    f = lambda *s: g.standard_normal(s).astype(np.float32)  # noqa: E731
    dim = ndirs * D_INNER
    dt0 = np.exp(g.random(dim) * (np.log(0.1) - np.log(1e-3)) + np.log(1e-3)).clip(min=1e-4)
    return dict(u=f(clips, dim, SEQLEN), delta=0.5 * f(clips, dim, SEQLEN), z=f(clips, D_INNER, SEQLEN),
                dout=f(clips, D_INNER, SEQLEN), B=f(clips, ndirs, D_STATE, SEQLEN), C=f(clips, ndirs, D_STATE, SEQLEN),
                A=-np.tile(np.arange(1, D_STATE + 1, dtype=np.float32), (dim, 1)), D=np.ones(dim, np.float32),
                bias=(dt0 + np.log(-np.expm1(-dt0))).astype(np.float32))


def cpu_block_step(t, ndirs):
    """fwd + bwd of the block's direction scans on the host: gather into traversal order, oracle.c, scatter back
    (oracle/dirs.py) -- the reference's own data flow (flip / interleave copies around the plain op)."""
    from oracle import dirs as odirs
    d = DIRS[:ndirs]
    odirs.scan_dirs_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["z"], t["bias"], True, d, NFRAMES)
    odirs.scan_dirs_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["z"], t["bias"], t["dout"], True, d, NFRAMES)


def time_cpu(steps, warmup, clips, ndirs):
    import oracle
    oracle.set_num_threads(os.cpu_count() or 1)
    t = block_inputs_numpy(clips, ndirs)
    for _ in range(warmup):
        cpu_block_step(t, ndirs)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_block_step(t, ndirs)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    fwd_b, bwd_b = algo_bytes(clips * ndirs)
    return (fwd_b + bwd_b) / dt / 1e9, dt, oracle.num_threads()


def time_torch_port():
    """The reference's own op-for-op CPU code (selective_scan_ref restated in oracle/torch_port.py) at the sizes SURVEY.md
    8(d) names: forward at the full length, forward + autograd backward at L = 4096 (its backward is O(L^2) and does not
    finish at 20480).  Nothing is extrapolated."""
    import torch
    from oracle.torch_port import selective_scan_port
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    res = {}
    for name, L, bwd in (("fwd_full_L_s", SEQLEN, False), ("fwd_bwd_L4096_s", 4096, True)):
        mk = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
        leaves = dict(u=mk(1, D_INNER, L), delta=0.5 * mk(1, D_INNER, L),
                      A=-torch.arange(1, D_STATE + 1, dtype=torch.float32).repeat(D_INNER, 1), B=mk(1, D_STATE, L),
                      C=mk(1, D_STATE, L), D=torch.ones(D_INNER), z=mk(1, D_INNER, L), bias=torch.rand(D_INNER, generator=g) - 4.0)
        if bwd:
            leaves = {k: v.requires_grad_() for k, v in leaves.items()}
        t0 = time.perf_counter()
        with torch.set_grad_enabled(bwd):
            out = selective_scan_port(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"],
                                      z=leaves["z"], delta_bias=leaves["bias"], delta_softplus=True)
            if bwd:
                out.backward(mk(1, D_INNER, L))
        res[name] = time.perf_counter() - t0
    return res


def cpu_baseline_obj(value, cores, clips, ndirs, extra=None):
    o = {"value": value, "unit": "GB/s", "cores": cores, "kind": "port",
         "sample": f"the full workload, one step: fwd+bwd of the {ndirs} direction scan(s) of {clips} clip(s), L = {SEQLEN}, "
                   f"d_inner {D_INNER}, d_state {D_STATE}; oracle/oracle.c (fp64 accumulation, O(L) analytic backward, "
                   f"OpenMP) behind the reference's gather / scatter data flow (oracle/dirs.py)"}
    if extra:
        o.update(extra)
    return o


def run_reference(args, rank, world):
    if rank != 0:
        return
    value, dt, cores = time_cpu(args.steps, args.warmup, args.clips, args.dirs)
    cfg, _ = workload_config(args.clips, args.dirs)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": cpu_baseline_obj(value, cores, args.clips, args.dirs),
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ScanSet:
    """One set of device-resident inputs + preallocated outputs/workspaces + prebuilt C-ABI args for the direction scans
    of one Temporal Mamba block.  ndirs = 3: channel-concatenated direction blocks, B / C as column blocks of an
    x_dbl-like (B, 3, L, R+2N) tensor (x_proj's GEMM output), dB / dC written into the matching columns of dx_dbl, z and
    dout shared by the directions.  ndirs = 1: the plain op, (B,1,N,L) B / C (the round-1 workload)."""

    def __init__(self, clips, device, seed, ndirs, flat=False):
        import torch
        from vivim_b200 import _lib
        g = torch.Generator(device="cpu").manual_seed(seed)
        bf = torch.bfloat16
        B_, D_, L_, N_, R_ = clips, D_INNER, SEQLEN, D_STATE, DT_RANK
        dim, R2 = ndirs * D_, R_ + 2 * N_
        U = (L_ + _lib.VV_SCAN_SEGMENT - 1) // _lib.VV_SCAN_SEGMENT
        r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
        self.host = dict(u=r(B_, dim, L_).to(bf), delta=(0.5 * r(B_, dim, L_)).to(bf), z=r(B_, D_, L_).to(bf),
                         dout=r(B_, D_, L_).to(bf))
        if ndirs > 1:
            self.host["x_dbl"] = r(B_, ndirs, L_, R2).to(bf)
        else:
            self.host.update(B=r(B_, 1, N_, L_).to(bf), C=r(B_, 1, N_, L_).to(bf))
        if flat:     # every input a view of one flat device buffer (one H2D copy per step in the e2e leg)
            self.offs, total = {}, 0
            for k, v in self.host.items():
                self.offs[k] = total
                total += -(-v.numel() * 2 // 256) * 256
            self.flat_in = torch.empty(total, dtype=torch.uint8, device=device)
            self.t = {k: self.flat_in[self.offs[k]:self.offs[k] + v.numel() * 2].view(bf).view(v.shape) for k, v in self.host.items()}
            for k, v in self.host.items():
                self.t[k].copy_(v)
        else:
            self.t = {k: v.to(device) for k, v in self.host.items()}
        # module init of mamba_simple.py:99-117: A = -(1..N), D = 1, dt bias = softplus^-1(U[1e-3, 1e-1])
        dt0 = torch.exp(torch.rand(dim, generator=g) * (torch.log(torch.tensor(0.1)) - torch.log(torch.tensor(1e-3)))
                        + torch.log(torch.tensor(1e-3))).clamp(min=1e-4)
        self.p = dict(A=-torch.arange(1, N_ + 1, dtype=torch.float32).repeat(dim, 1).to(device),
                      D=torch.ones(dim, device=device), bias=(dt0 + torch.log(-torch.expm1(-dt0))).to(device))
        # outputs: one flat buffer (one D2H copy per step in the e2e leg)
        n_act = B_ * dim * L_
        n_bc_io = B_ * ndirs * L_ * R2 if ndirs > 1 else 2 * B_ * N_ * L_
        self.flat_out = torch.empty((4 * n_act + n_bc_io) * 2, dtype=torch.uint8, device=device)
        fo = self.flat_out.view(bf)
        self.out_z, self.du, self.ddelta, self.dz = (fo[i * n_act:(i + 1) * n_act].view(B_, dim, L_) for i in range(4))
        n_bc = B_ * ndirs * N_ * L_
        self.acc = torch.zeros(2 * n_bc + dim * N_ + 2 * dim, dtype=torch.float32, device=device)
        self.agg = torch.empty(B_, dim, U, N_, 2, dtype=torch.float32, device=device)
        self.chk = torch.empty(B_, dim, U, N_, dtype=torch.float32, device=device)
        self.radj = torch.empty(B_, dim, U, N_, dtype=torch.float32, device=device)
        self.last = torch.empty(B_, dim, N_, dtype=torch.float32, device=device)
        a = _lib.ScanArgs()
        t, p = self.t, self.p
        a.u, a.delta, a.z, a.dout = (t[k].data_ptr() for k in ("u", "delta", "z", "dout"))
        a.A, a.D, a.delta_bias = p["A"].data_ptr(), p["D"].data_ptr(), p["bias"].data_ptr()
        a.out_z, a.du, a.ddelta, a.dz = (x.data_ptr() for x in (self.out_z, self.du, self.ddelta, self.dz))
        a.last_state, a.agg, a.chk, a.radj = (x.data_ptr() for x in (self.last, self.agg, self.chk, self.radj))
        base = self.acc.data_ptr()
        a.dB, a.dC = base, base + 4 * n_bc
        a.dA = base + 8 * n_bc
        a.dD = a.dA + 4 * dim * N_
        a.ddelta_bias = a.dD + 4 * dim
        a.batch, a.dim, a.seqlen, a.dstate, a.ngroups = B_, dim, L_, N_, ndirs
        for name in ("u", "delta", "outz", "du", "ddelta", "dz"):
            setattr(a, name + "_bs", dim * L_)
            setattr(a, name + "_ds", L_)
        a.z_bs = a.dout_bs = D_ * L_
        a.z_ds = a.dout_ds = L_
        a.A_ds, a.A_ns = N_, 1
        if ndirs > 1:
            xd = t["x_dbl"]
            self.dx_dbl = fo[4 * n_act:].view(B_, ndirs, L_, R2)
            esz = 2
            a.Bm, a.Cm = xd.data_ptr() + R_ * esz, xd.data_ptr() + (R_ + N_) * esz
            a.dB_io, a.dC_io = self.dx_dbl.data_ptr() + R_ * esz, self.dx_dbl.data_ptr() + (R_ + N_) * esz
            for pre in ("B_", "C_", "dBio_", "dCio_"):
                setattr(a, pre + "bs", ndirs * L_ * R2)
                setattr(a, pre + "gs", L_ * R2)
                setattr(a, pre + "ns", 1)
                setattr(a, pre + "ls", R2)
            a.ndirs, a.nframes, a.gate_rows = ndirs, NFRAMES, D_
            for k, m in enumerate(DIRS[:ndirs]):
                a.dir_mode[k] = {"fwd": _lib.VV_DIR_FWD, "rev": _lib.VV_DIR_REV, "frames": _lib.VV_DIR_FRAMES}[m]
        else:
            self.dBC16 = fo[4 * n_act:].view(2, B_, 1, N_, L_)
            a.Bm, a.Cm = t["B"].data_ptr(), t["C"].data_ptr()
            a.B_bs = a.C_bs = a.B_gs = a.C_gs = N_ * L_
            a.B_ns = a.C_ns = L_
            a.dB_io, a.dC_io = self.dBC16[0].data_ptr(), self.dBC16[1].data_ptr()
        a.io_dtype, a.delta_softplus, a.zero_accumulators = _lib.VV_BF16, 1, 1
        self.args = a
        self.ndirs = ndirs

    def input_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.t.values())


def launch_step(s, lib, stream):
    """scan fwd (3 kernels) -> scan bwd (3 kernels, the first one zero-fills the fp32 accumulators) -> dB/dC cast into the
    I/O dtype / into dx_dbl (a fourth kernel of vv_scan_bwd, chained with programmatic dependent launch)."""
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


def graphs_for(sets, lib, device, torch):
    side = torch.cuda.Stream(device)
    with torch.cuda.stream(side):
        for s in sets:                      # warm-up outside capture (sets func attributes, fills workspaces)
            launch_step(s, lib, side.cuda_stream)
    side.synchronize()
    graphs = []
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            launch_step(s, lib, torch.cuda.current_stream().cuda_stream)
        graphs.append(g)
    torch.cuda.synchronize()
    return graphs


def teardown(dist, world):
    """Leave the process group without ever hanging the job: the result line is already printed and flushed."""
    if world <= 1:
        return
    import threading as _th
    _th.Timer(30.0, lambda: os._exit(0)).start()     # a communicator teardown that blocks must not keep the launcher alive
    try:
        import torch
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    finally:
        os._exit(0)


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
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        dist.init_process_group("nccl", device_id=device)

    if args.only_vivim:      # development: the whole-network legs alone (not a bench line)
        v = bench_vivim(args, rank, world, device, torch, dist)
        v.pop("_launches", None)
        if rank == 0:
            print(json.dumps({"n_gpus": world, **v}), flush=True)
        teardown(dist, world)
        return
    clips, ndirs = args.clips, args.dirs
    fwd_b, bwd_b = algo_bytes(clips * ndirs)
    cfg, n_sets = workload_config(clips, ndirs)
    # inputs larger than L2: rotate over enough input sets that a set is evicted before it is reused
    sets = [ScanSet(clips, device, seed=1000 * rank + i, ndirs=ndirs) for i in range(n_sets)]
    graphs = graphs_for(sets, lib, device, torch)

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
    timed_launches = 7 * args.steps       # per rank: 3 forward + 4 backward kernels of libvivim_b200.so per timed step
    launches = timed_launches

    # ---- the six kernels one by one (pass mask), rotating sets, CUDA events on the launch stream
    stream = torch.cuda.current_stream().cuda_stream
    passes = {}
    reps = max(10, min(50, args.steps))
    for bwd in (0, 1):
        fn = lib.vv_scan_bwd if bwd else lib.vv_scan_fwd
        for bit, name in ((1, "agg"), (2, "carry"), (4, "main")) + (((8, "cast"),) if bwd else ()):
            for s in sets:
                s.args.pass_mask = bit
            run = lambda i: _lib.check(fn(ctypes.byref(sets[i % n_sets].args), ctypes.c_void_p(stream)), "scan pass")  # noqa: E731
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            passes[("bwd_" if bwd else "fwd_") + name] = time_events(run, reps, torch) / reps
    for s in sets:
        s.args.pass_mask = 0
    launches += 7 * (reps + 3)
    t_main = passes["bwd_main"]
    peak, peak_src = measured_peak()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("seg_bwd_kernel_bytes_per_launch_block" if ndirs > 1 else "seg_bwd_kernel_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "seg_bwd_kernel", "achieved": bwd_b / t_main / 1e9, "peak": peak,
                "unit": "GB/s", "frac": bwd_b / t_main / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bwd_b, "us_per_launch": t_main * 1e6,
                "note": "N=16 states per (channel, token): ~26 issue slots, 1 MUFU and 6.6 shared-memory/shuffle wavefront "
                        "bytes per state-step make this kernel issue / LSU bound, not HBM bound, on B200 (DESIGN.md section 5)"}

    # ---- the whole block in one launch chain: the three direction scans of Mamba.forward v3 as direction blocks of ONE
    #      vv_scan_fwd / vv_scan_bwd (B / C from x_dbl rows, z / dout shared), what vivim_b200.mamba_block runs
    block = None
    if ndirs == 1 and not args.quick:
        nb = len(DIRS)
        sb = [ScanSet(clips, device, seed=77 + i, ndirs=nb) for i in range(workload_config(clips, nb)[1])]
        gb = graphs_for(sb, lib, device, torch)
        for i in range(10):
            gb[i % len(gb)].replay()
        torch.cuda.synchronize()
        n1 = max(50, min(args.steps, 1000))
        tb = max_over_ranks(time_events(lambda i: gb[i % len(gb)].replay(), n1, torch), device) / n1
        fb, bb = algo_bytes(clips * nb)
        kb = {}
        for bwd in (0, 1):
            fn = lib.vv_scan_bwd if bwd else lib.vv_scan_fwd
            for bit, name in ((1, "agg"), (2, "carry"), (4, "main")) + (((8, "cast"),) if bwd else ()):
                for s_ in sb:
                    s_.args.pass_mask = bit
                run = lambda i: _lib.check(fn(ctypes.byref(sb[i % len(sb)].args), ctypes.c_void_p(stream)), "scan pass")  # noqa: E731
                for i in range(3):
                    run(i)
                torch.cuda.synchronize()
                kb[("bwd_" if bwd else "fwd_") + name] = time_events(run, reps, torch) / reps * 1e6
        block = {"directions": list(DIRS), "us_per_block": tb * 1e6, "us_per_direction_scan": tb * 1e6 / (clips * nb),
                 "GBps": (fb + bb) / tb / 1e9, "kernel_us": kb,
                 "bytes_algorithmic": fb + bb, "bytes_compulsory": sum(compulsory_bytes(clips, nb)),
                 "note": "the three direction scans of one Temporal Mamba block (mamba_simple.py:217-260) as ONE launch chain: "
                         "direction blocks with per-block traversal order, B / C read from x_dbl rows, dB / dC written into dx_dbl, z "
                         "and dout shared; GBps counts the reference's three calls (3 x (11T + 6S))"}
        launches += 7 * (n1 + 10 + len(gb)) + 7 * (reps + 3)
        del sb, gb

    # ---- conv1d at the same shape: all directions in one launch, and the single-direction kernels
    conv = bench_conv(clips, len(DIRS) if ndirs == 1 else ndirs, device, torch)
    launches += conv.pop("_launches")

    # ---- end to end: plugin boundary, host (pinned) buffers in, all results out, every step
    e2e_steps = max(4, min(args.steps, 40))
    e2e = bench_e2e(clips, ndirs, e2e_steps, world, device, torch, dist, lib)
    launches += 7 * (3 * e2e_steps + 4)

    # ---- the unmodified reference kernels on this GPU (baseline/_ref), for the record
    ref_cuda = None if args.quick else bench_ref_cuda(clips, device, torch)

    # ---- second half of the metric: Vivim clips/s (training step, inference) on this many GPUs
    vivim = {}
    if not args.no_vivim:
        vivim = bench_vivim(args, rank, world, device, torch, dist)
        launches += vivim.pop("_launches", 0)

    if rank != 0:
        teardown(dist, world)
        return
    # Second roofline, the one that actually binds these kernels: MUFU.EX2 lane-operations per pass (one per state-step
    # for the decay, plus softplus / sigmoid in the per-position pre-pass) against the measured MUFU rate
    # (scripts/microbench/mufu_rate.cu: 15.57 lanes/clk/SM on B200) at the SM clock sampled during the run.
    ss = clips * ndirs * D_INNER * SEQLEN * D_STATE            # state-steps
    pos = clips * ndirs * D_INNER * SEQLEN                     # (channel, position) pairs
    mufu_ops = {"fwd_agg": ss + 2 * pos, "fwd_main": ss + 2 * pos + 2 * pos, "bwd_agg": ss + 4 * pos,
                "bwd_main": ss * 9 // 8 + 5 * pos}
    sm_mhz = (clk.summary().get("sm_mhz") or 1965.0)
    mufu_peak = 15.57 * 148 * sm_mhz * 1e6
    mufu = {k: {"mufu_lane_ops": v, "floor_us": v / mufu_peak * 1e6, "frac_of_mufu_peak": v / mufu_peak / passes[k]}
            for k, v in mufu_ops.items()}
    cf, cb = compulsory_bytes(clips, ndirs)
    line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
            "us_per_direction_scan": elapsed / args.steps / (clips * ndirs) * 1e6,
            "bytes_algorithmic_per_step": fwd_b + bwd_b, "bytes_compulsory_per_step": cf + cb,
            "hbm_frac_of_measured_peak": value / world / peak,
            "roofline": roofline, "e2e": e2e, "gpu_launches": timed_launches, "gpu_launches_whole_run": launches,
            "clocks": clk.summary(),
            "kernel_us": {k: v * 1e6 for k, v in passes.items()},
            "walking_passes": {"fwd": 2, "bwd": 2,
                               "note": "decays are evaluated once per (channel, token, state) in each of: forward aggregate, "
                                       "forward main, reverse aggregate, backward main (DESIGN.md section 5 says why not fewer)"},
            "mufu_roofline": {"peak_lane_ops_per_s": mufu_peak, "source": "scripts/microbench/mufu_rate.cu (15.57 lanes/clk/SM)",
                              "kernels": mufu},
            "fwd_GBps_kernels_only": fwd_b / (passes["fwd_agg"] + passes["fwd_carry"] + passes["fwd_main"]) / 1e9,
            "bwd_GBps_kernels_only": bwd_b / (passes["bwd_agg"] + passes["bwd_carry"] + passes["bwd_main"] + passes["bwd_cast"]) / 1e9,
            "conv1d": conv}
    if block:
        line["block_three_directions"] = block
    if ref_cuda:
        line["ref_cuda_us"] = ref_cuda
    line.update(vivim)
    if world == 1:
        cpu_v, _, cores = time_cpu(steps=2, warmup=1, clips=clips, ndirs=ndirs)
        line["cpu_baseline"] = cpu_baseline_obj(cpu_v, cores, clips, ndirs,
                                                None if args.quick else {"torch_port_selective_scan_ref": time_torch_port()})
    print(json.dumps(line), flush=True)
    sys.stdout.flush()
    teardown(dist, world)


def bench_conv(clips, ndirs, device, torch):
    from vivim_b200 import causal_conv1d_cuda as ccc
    bf = torch.bfloat16
    T = clips * D_INNER * SEQLEN * 2
    n_sets = max(2, -(-2 * L2_BYTES // ((1 + ndirs) * T)) + 1)
    xz = [torch.randn(clips, 2 * D_INNER, SEQLEN, device=device, dtype=bf) for _ in range(n_sets)]
    dout = [torch.randn(clips, ndirs * D_INNER, SEQLEN, device=device, dtype=bf) for _ in range(n_sets)]
    w = torch.randn(ndirs, D_INNER, 4, device=device)
    b = torch.randn(ndirs, D_INNER, device=device)
    dxz = torch.empty_like(xz[0])
    reps = 30
    d = DIRS[:ndirs]
    cases = [("fwd", lambda i: ccc.causal_conv1d_fwd(xz[i % n_sets][:, :D_INNER], w[0], b[0], True), 2 * T),
             ("bwd", lambda i: ccc.causal_conv1d_bwd(xz[i % n_sets][:, :D_INNER], w[0], b[0], dout[i % n_sets][:, :D_INNER],
                                                     dxz[:, :D_INNER], True), 3 * T)]
    if ndirs > 1:
        # algorithmic bytes of the reference's ndirs separate calls: 2T / 3T each; this launch moves (1+ndirs)T both ways
        cases += [("dirs_fwd", lambda i: ccc.causal_conv1d_dirs_fwd(xz[i % n_sets][:, :D_INNER], w, b, d, NFRAMES, True), ndirs * 2 * T),
                  ("dirs_bwd", lambda i: ccc.causal_conv1d_dirs_bwd(xz[i % n_sets][:, :D_INNER], w, b, dout[i % n_sets],
                                                                    dxz[:, :D_INNER], d, NFRAMES, True), ndirs * 3 * T)]
    res = {}
    for name, fn, nbytes in cases:
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
    if ndirs > 1:
        res["dirs_fwd_moved_GBps"] = (1 + ndirs) * T / (res["dirs_fwd_us"] * 1e-6) / 1e9
        res["dirs_bwd_moved_GBps"] = (2 + ndirs) * T / (res["dirs_bwd_us"] * 1e-6) / 1e9
    res["note"] = (f"x = first half of xz ({clips},{2 * D_INNER},{SEQLEN}) bf16, K=4, SiLU; fwd/bwd: one direction (2T / 3T); "
                   f"dirs_*: the {ndirs} directions in one launch, GBps against the reference's {ndirs} calls (2T / 3T each), "
                   f"*_moved_GBps against what the launch itself reads and writes")
    res["_launches"] = len(cases) * n_sets * (reps + 3)
    return res


def bench_e2e(clips, ndirs, steps, world, device, torch, dist, lib):
    """The block's scans through the plugin boundary (vivim_b200.selective_scan_cuda -- the module standing where the
    reference's pybind `selective_scan_cuda` stands -- i.e. vv_scan_fwd / vv_scan_bwd on caller-owned device buffers),
    with HOST buffers: every step copies ALL inputs in from pinned host memory (one flat H2D copy) and ALL results
    (out_z, du, ddelta, dz, dB / dC) back out (one flat D2H copy).  Three streams, two buffer sets: the copies of steps
    i+1 and i-1 overlap the kernels of step i."""
    from vivim_b200.sharding import aggregate_throughput, max_over_ranks
    depth = 2
    sets = [ScanSet(clips, device, seed=500 + k, ndirs=ndirs, flat=True) for k in range(depth)]
    pin_in = torch.empty(sets[0].flat_in.numel(), dtype=torch.uint8, pin_memory=True)
    pin_in.copy_(sets[0].flat_in)
    host_out = [torch.empty(sets[0].flat_out.numel(), dtype=torch.uint8, pin_memory=True) for _ in range(depth)]
    h2d, d2h = sets[0].input_bytes(), sets[0].flat_out.numel()
    s_in, s_cmp, s_out = (torch.cuda.Stream(device) for _ in range(3))
    ev_in = [torch.cuda.Event() for _ in range(depth)]
    ev_cmp = [torch.cuda.Event() for _ in range(depth)]
    ev_out = [torch.cuda.Event() for _ in range(depth)]

    kernels_on = [True]

    def one(i):
        k = i % depth
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp[k])              # the kernels that last read this input set are done
            sets[k].flat_in.copy_(pin_in, non_blocking=True)
            ev_in[k].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in[k])
            s_cmp.wait_event(ev_out[k])             # the previous results of this slot have left the device
            if kernels_on[0]:
                launch_step(sets[k], lib, s_cmp.cuda_stream)
            ev_cmp[k].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp[k])
            host_out[k].copy_(sets[k].flat_out, non_blocking=True)
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
    elapsed = max_over_ranks(blocks[1], device)
    # the same copies WITHOUT the kernels: what the host link of this box gives `world` ranks at once (attributes the
    # e2e figure to PCIe / host memory when N ranks share one host)
    kernels_on[0] = False
    copies_only = max_over_ranks(sorted(block() for _ in range(3))[1], device)
    kernels_on[0] = True
    fwd_b, bwd_b = algo_bytes(clips * ndirs)
    return {"value": aggregate_throughput((fwd_b + bwd_b) * steps, world, elapsed) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": steps, "ms_per_step": elapsed / steps * 1e3,
            "ms_per_step_blocks": [b / steps * 1e3 for b in blocks],
            "copies_only_ms_per_step": copies_only / steps * 1e3,
            "copies_only_GBps_per_rank_each_way": max(h2d, d2h) / (copies_only / steps) / 1e9,
            "pcie_floor_ms": max(h2d, d2h) / 55e9 * 1e3,
            "api": "vivim_b200.selective_scan_cuda fwd + bwd (vv_scan_fwd / vv_scan_bwd, include/vivim_b200.h) on caller-owned "
                   "device buffers, inputs from / results to pinned host memory every step",
            "pipeline": "one flat H2D copy / kernels / one flat D2H copy on three streams, two buffer sets: copies of "
                        "step i+1 and i-1 overlap the kernels of step i"}


def bench_ref_cuda(clips, device, torch):
    """fwd + bwd of the reference's own CUDA kernels (selective_scan_cuda / causal_conv1d_cuda built unmodified for
    sm_100a by baseline/build_ref.py) at the stage-1 shape, one direction scan, same CUDA-event harness.  A note beside
    our numbers, not a target; absent when baseline/_ref was not built."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "selective_scan_cuda.so")):
        return None
    try:
        sys.path.insert(0, ref_dir)
        import causal_conv1d_cuda as rconv
        import selective_scan_cuda as rscan
    except Exception as e:   # noqa: BLE001
        return {"unavailable": str(e)[:200]}
    finally:
        sys.path.remove(ref_dir)
    bf = torch.bfloat16
    B_, D_, L_, N_ = clips, D_INNER, SEQLEN, D_STATE
    n_sets = workload_config(clips, 1)[1]
    mk = lambda *s: torch.randn(*s, device=device).to(bf)  # noqa: E731
    sets = [dict(u=mk(B_, D_, L_), delta=(0.5 * torch.randn(B_, D_, L_, device=device)).to(bf), z=mk(B_, D_, L_),
                 B=mk(B_, 1, N_, L_), C=mk(B_, 1, N_, L_), dout=mk(B_, D_, L_)) for _ in range(n_sets)]
    A = -torch.arange(1, N_ + 1, dtype=torch.float32, device=device).repeat(D_, 1)
    Dv, bias = torch.ones(D_, device=device), torch.rand(D_, device=device) - 4.6
    w, cb = torch.randn(D_, 4, device=device), torch.randn(D_, device=device)

    def scan(i):
        s = sets[i % n_sets]
        out, x, out_z = rscan.fwd(s["u"], s["delta"], A, s["B"], s["C"], Dv, s["z"], bias, True)
        rscan.bwd(s["u"], s["delta"], A, s["B"], s["C"], Dv, s["z"], bias, s["dout"], x, out, None, True, False)

    def conv(i):
        s = sets[i % n_sets]
        rconv.causal_conv1d_fwd(s["u"], w, cb, True)
        rconv.causal_conv1d_bwd(s["u"], w, cb, s["dout"], None, True)

    res = {}
    for name, fn in (("scan_fwd_bwd", scan), ("conv_fwd_bwd", conv)):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        reps = 30
        res[name] = time_events(fn, reps, torch) / reps * 1e6
    res["note"] = ("unmodified reference kernels compiled for sm_100a, one direction scan (B=%d, D=%d, L=%d, N=%d, bf16), eager "
                   "launches through the reference's pybind modules, including their output allocations" % (B_, D_, L_, N_))
    return res


def bench_vivim(args, rank, world, device, torch, dist):
    """configs[2] / configs[3]: Vivim clips/s.  Whole network (SegFormer-b3 stages interleaved with Temporal Mamba
    stages, decode head; vivim_b200/temporal_model.py restates modeling/vivim.py), synthetic clips 5 x 3 x 256 x 256, 3
    classes, random init, bf16 autocast.  Training: batch 3 per GPU, recall_focused_loss + AdamW(lr 1e-4, wd 1e-2)
    (multiclass_training_folds.py:339-361, 505), forward + backward replayed as ONE CUDA graph, gradients in one flat
    buffer all-reduced over NCCL (N > 1), fused AdamW.  Inference: batch 4 per GPU, the forward as one CUDA graph."""
    from vivim_b200.graphed import InferenceGraph, TrainStepGraph
    from vivim_b200.temporal_model import RecallFocusedLoss, Vivim
    res = {}
    launches = 0
    frames, image = NFRAMES, 256
    steps = max(5, min(args.steps, 20))
    warm = 3

    def timed(step_fn, batch):
        # clocks of this leg too: it runs after a minute of scan benchmarks, on a warm GPU
        with ClockSampler(device.index if device.index is not None else 0) as clk:
            for _ in range(warm):
                step_fn()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            clk.wait_first_sample()
            clk.mark()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = step_fn()
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * batch * steps / t.item(), t.item() / steps * 1e3, float(out), clk.summary()

    torch.manual_seed(0)
    model = Vivim(out_chans=3).to(device)
    for name, p in model.named_parameters():        # unused by the forward, as in the reference
        if "downsample_layers.layer_norm" in name or "decoder.classifier" in name:
            p.requires_grad_(False)
    n_ln = 0
    if not args.torch_layernorm:
        # the LayerNorm kernels written for the Temporal Mamba blocks also serve the SegFormer stages' nn.LayerNorm modules
        from vivim_b200.layernorm import use_token_layernorm
        n_ln = use_token_layernorm(model)
    # ---- training (configs[2])
    batch = 3
    clip = torch.randn(batch, frames, 3, image, image, device=device)
    target = torch.randint(0, 3, (batch * frames, image, image), device=device)
    model.train()
    loss_fn = RecallFocusedLoss().to(device)
    overlapped = world > 1 and args.overlap_allreduce
    tsg = TrainStepGraph(model, loss_fn, (clip,), (target,), autocast_dtype=torch.bfloat16,
                         grad_allreduce=(lambda t: dist.all_reduce(t, op=dist.ReduceOp.AVG)) if overlapped else None)
    opt = torch.optim.AdamW(tsg.params, lr=1e-4, weight_decay=1e-2, betas=(0.9, 0.999), fused=True, capturable=True)

    def train_step():
        loss = tsg()
        if world > 1 and not overlapped:
            dist.all_reduce(tsg.flat_grad, op=dist.ReduceOp.AVG)
        opt.step()
        return loss

    cps, ms, loss, clocks = timed(train_step, batch)
    res.update(vivim_train_clips_per_s=cps, vivim_train_ms_per_step=ms,
               vivim_train={"batch_per_gpu": batch, "steps": steps, "loss_after": loss, "loss": "recall_focused_loss",
                            "optimizer": "AdamW lr 1e-4 wd 1e-2 (fused)",
                            "grad_allreduce": ("none (1 GPU)" if world == 1 else
                                               ("%d NCCL all_reduce(AVG) buckets of the flat fp32 gradient (%.0f MB) captured INSIDE the "
                                                "graph on a side stream, each launched when the backward has produced its last gradient"
                                                % (len(tsg._buckets), tsg.flat_grad.numel() * 4 / 1e6)) if overlapped else
                                               "one NCCL all_reduce(AVG) of the flat fp32 gradient (%.0f MB) after the graph replay"
                                               % (tsg.flat_grad.numel() * 4 / 1e6)),
                            "launch": "forward + backward as one CUDA graph (vivim_b200.graphed.TrainStepGraph)",
                            "clocks": clocks})
    tsg.close()         # the graph holds NCCL nodes: release it before the process group is torn down
    del tsg, opt
    # ---- inference (configs[3]: 32 clips over 8 GPUs -> 4 per GPU)
    batch = 4
    model.eval()
    clip = torch.randn(batch, frames, 3, image, image, device=device)
    infer = InferenceGraph(model, (clip,), autocast_dtype=torch.bfloat16)
    cps, ms, _, clocks = timed(lambda: infer().float().mean(), batch)
    res.update(vivim_infer_clips_per_s=cps, vivim_infer_ms_per_step=ms,
               vivim_infer={"batch_per_gpu": batch, "steps": steps, "launch": "forward as one CUDA graph (InferenceGraph)",
                            "clocks": clocks})
    res["vivim_config"] = {"workload": "Vivim multiclass, image 256, clip_length 5, 3 classes, random init, synthetic clips, bf16 autocast",
                           "layernorm": ("vivim_b200 TokenLayerNorm in the Temporal Mamba blocks and, re-classed in place, in %d SegFormer "
                                         "nn.LayerNorm modules" % n_ln) if n_ln else "vivim_b200 TokenLayerNorm in the Temporal Mamba blocks only",
                           "model": "vivim_b200.temporal_model.Vivim (restates modeling/vivim.py; state-dict compatible)"}
    res["_launches"] = launches
    infer.graph.reset()
    del infer, model
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--clips", type=int, default=1, help="clips per GPU (BASELINE configs[1]: 1)")
    ap.add_argument("--dirs", type=int, default=1, choices=(1, 2, 3),
                    help="direction scans per launch chain of the headline workload (1 = BASELINE configs[1], one scan over "
                         "d_inner 128; 3 = the block's three directions fused, also reported as block_three_directions)")
    ap.add_argument("--dir-modes", default=None, help="kernel development: comma list overriding the traversal order of the "
                                                      "direction blocks, e.g. fwd,fwd,fwd")
    ap.add_argument("--no-vivim", action="store_true", help="skip the whole-network clips/s legs")
    ap.add_argument("--only-vivim", action="store_true", help="development: run only the whole-network legs")
    ap.add_argument("--torch-layernorm", action="store_true", help="keep torch's LayerNorm in the SegFormer stages")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="training: bucketed all-reduces captured inside the step graph on a side stream (TrainStepGraph "
                         "grad_allreduce) instead of one all-reduce after the replay.  Measured on 8 B200: 44.77 ms vs 44.39 ms "
                         "per step -- the NCCL CTAs take more from the backward than the 3 ms collective they hide -- so off by default")
    ap.add_argument("--no-overlap", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--quick", action="store_true", help="kernel development: skip the slow side measurements")
    args = ap.parse_args()
    if args.dir_modes:
        global DIRS
        DIRS = tuple(args.dir_modes.split(","))
        args.dirs = len(DIRS)
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
