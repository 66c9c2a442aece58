#!/usr/bin/env python
"""bench.py — LazySum mul! throughput (amplitude-updates/s and HBM GB/s), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

N = 1: BASELINE config 4 — Heisenberg spin chain N=28 (periodic, 84 LazyTensor terms with sparse sigma
factors exactly as src/spin.jl builds them), LazySum mul! on a Ket of 2^28 ComplexF64 amplitudes (4 GiB).
N > 1: BASELINE config 5 — the same chain at N=33 (99 terms) sharded over the ranks (one process per GPU,
launched by torchrun), axis swaps over NCCL; if 4 slabs of 2^(33-p) amplitudes do not fit one GPU the largest
chain that does is used and named in config.workload.

One "step" = one complete mul!(y, H, x, alpha, 0).  One amplitude-update = one output amplitude of one
complete mul! (all terms accumulated) — SURVEY.md §8(d).  Timing: CUDA events on the launching stream,
barrier + synchronize on both sides, max over ranks; the 4 GiB+ operands exceed the 126 MB L2 so every step
streams from HBM (no flush needed; stated in config).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LazySum mul! amplitude-updates/s"
UNIT = "amplitude-updates/s"


def chain_spec(n, seed=2024):
    """Heisenberg XYZ chain, periodic: 3N terms J^a_i sigma^a_i sigma^a_{i+1}; coefficients uniform[0.5,1.5)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    spec = []
    for i in range(1, n + 1):
        j = i % n + 1
        for a in range(3):
            spec.append((float(rng.uniform(0.5, 1.5)), sorted([i, j]), a))
    return spec


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML from a background thread (a sample every
    ~2 ms: the timed region of the default run is well under 100 ms, shorter than one `nvidia-smi -lms` period);
    `nvidia-smi` polling is the fallback when the NVML binding is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.p = self.f = self.thread = None
        self.sm, self.mx, self.bits = [], 0.0, 0
        try:
            import threading

            import pynvml

            pynvml.nvmlInit()
            handle = None
            try:
                import torch

                pr = torch.cuda.get_device_properties(gpu_index)
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self._stop = threading.Event()

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                        self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                    except Exception:
                        pass
                    self._stop.wait(0.002)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if self.sm:
                out = {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.mx,
                       "reasons": sorted(nm for bit, nm in self.REASONS.items() if self.bits & bit), "samples": len(self.sm),
                       "source": "nvml"}
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                   "source": "nvidia-smi"}
        return out


# --------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU algorithm for this workload: per-term passes through the pure-sparse recursion
    (src/operators_lazysum.jl:189-200 -> src/operators_lazytensor.jl:652-685), restated in oracle/qob_oracle.c
    (Julia is not installed in this image, so the reference itself cannot run: kind = "port").  The reference has
    no threading on this path, so 1 thread is all it can use."""
    if rank != 0:
        return
    import numpy as np
    import scipy.sparse as sp

    from oracle import qob_oracle as O

    # the same chain the product arm runs: N=28 on one GPU; sharded, the largest that fits three slabs per GPU of 180 GB
    # (N=33 from 4 GPUs on, N=32 on 2)
    n = 28 if world == 1 else (32 if world == 2 else 33)
    nterms = 3 * n
    total = args.steps + args.warmup
    pa = [np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]], dtype=complex),
          np.array([[1, 0], [0, -1]], dtype=complex)]
    # one step = ONE term of the sum applied to a 2^m-amplitude chain with the same bond structure; m is chosen from
    # a quick calibration (one term at 2^20) so that the whole --steps/--warmup run stays within ~2 minutes
    cal = (2,) * 20
    lt = O.LazyTensor(cal, cal, [10, 11], [O.Op((2,), (2,), sp.csc_matrix(pa[0]))] * 2)
    xc, yc = O.Ket(cal, O.fill_state(1 << 20, 1, 1e-3)), O.Ket(cal, np.zeros(1 << 20, dtype=complex))
    t0 = time.perf_counter()
    O.mul(yc, lt, xc, 1.0, 1.0)
    per_amp = (time.perf_counter() - t0) / (1 << 20)
    budget = 120.0 / max(total, 1)
    m = min(28, n)
    while m > 20 and per_amp * (1 << m) * 1.3 > budget:
        m -= 1
    dims = (2,) * m
    spec = chain_spec(m)
    x = O.Ket(dims, O.fill_state(1 << m, 7, 2.0 ** (-m / 2)))
    y = O.Ket(dims, np.zeros(1 << m, dtype=complex))
    times = []
    for s in range(total):
        c, idx, a = spec[(s * 7) % len(spec)]
        op = [O.Op((2,), (2,), sp.csc_matrix(pa[a]))] * 2
        lt = O.LazyTensor(dims, dims, idx, op)
        t0 = time.perf_counter()
        O.mul(y, lt, x, c, 1.0)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    t_term = sum(times) / len(times)
    # a complete mul! = nterms such passes; amplitude-updates/s is size-independent for this streaming algorithm
    value = (1 << m) / (nterms * t_term)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * nterms * t_term * (2.0 ** (n - m)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Heisenberg XYZ spin-1/2 chain N={n} periodic, LazySum of {nterms} LazyTensor terms, mul! on Ket",
                   "note": "ms_per_step extrapolated to the full workload from the timed sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{len(times)} single-term passes on a 2^{m}-amplitude state (of {nterms} terms on 2^{n}); "
                                   f"{t_term:.3f} s/term; reference path has no threading"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(n, nterms, seconds=20.0):
    """cpu_baseline leg of the default run: the oracle port timed on a bounded sample (rank 0, N=1 only)."""
    import numpy as np
    import scipy.sparse as sp

    from oracle import qob_oracle as O

    m = min(n, 26)
    dims = (2,) * m
    sx = sp.csc_matrix(np.array([[0, 1], [1, 0]], dtype=complex))
    sz = sp.csc_matrix(np.array([[1, 0], [0, -1]], dtype=complex))
    x = O.Ket(dims, O.fill_state(1 << m, 7, 2.0 ** (-m / 2)))
    y = O.Ket(dims, np.zeros(1 << m, dtype=complex))
    cases = [([1, 2], sx), ([m // 2, m // 2 + 1], sx), ([m - 1, m], sx), ([1, m], sz)]
    times = []
    t_start = time.perf_counter()
    for idx, s in cases:
        lt = O.LazyTensor(dims, dims, idx, [O.Op((2,), (2,), s)] * 2)
        t0 = time.perf_counter()
        O.mul(y, lt, x, 0.5, 1.0)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > seconds:
            break
    t_term = sum(times) / len(times)
    return {"value": (1 << m) / (nterms * t_term), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{len(times)} single-term passes of the restated sparse recursion (operators_lazytensor.jl:652-685) "
                      f"on a 2^{m}-amplitude state, {t_term:.3f} s/term, scaled by {nterms} terms per mul!"}


# --------------------------------------------------------------------------------------- product arm
def build_chain(Q, n):
    b = Q.SpinBasis(0.5)
    B = Q.tensor(*[b] * n)
    sig = (Q.sigmax(b), Q.sigmay(b), Q.sigmaz(b))
    terms, coefs = [], []
    for c, idx, a in chain_spec(n):
        terms.append(Q.LazyTensor(B, idx, (sig[a], sig[a])))
        coefs.append(c)
    return B, Q.LazySum(coefs, terms)


def run_product(args, rank, world, local_rank):
    import torch

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    import qob200 as Q

    warmup = max(args.warmup, 3)
    steps = args.steps
    alpha = 0.5 - 1.0j
    peak, peak_src = measured_peaks()

    n = args.spins or 28
    B, H = build_chain(Q, n)
    nloc = n
    x = Q.Ket(B)
    Q.fill_state(x.data, 7, 2.0 ** (-n / 2))
    y = Q.Ket(B)
    plan = Q.describe(H)

    def step():
        Q.mul_(y, H, x, alpha, 0.0)
    xs, ys = x.data, y.data
    run_common(args, Q, rank, world, step, xs, ys, n, nloc, plan, warmup, steps, peak, peak_src, H, alpha)


DEFAULT_EXCHANGE = "capi"   # the sharded apply behind the C ABI (qob_dist_*), direct mode; "fused": the Python orchestration over
                            # torch symmetric memory; "nccl": all-to-all axis swaps


def host_memory_available():
    """bytes of host memory this process tree may still take: min(MemAvailable, cgroup limit - cgroup usage)"""
    avail = float("inf")
    try:
        import psutil

        avail = float(psutil.virtual_memory().available)
    except Exception:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        if lim != "max":
            avail = min(avail, float(lim) - float(open("/sys/fs/cgroup/memory.current").read().strip()))
    except Exception:
        pass
    return avail


def run_product_dist(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    import qob200 as Q
    from qob200.dist import ShardedLazySum

    warmup = max(args.warmup, 3)
    alpha = 0.5 - 1.0j
    peak, peak_src = measured_peaks()
    p = world.bit_length() - 1
    free, _total = torch.cuda.mem_get_info()
    n = args.spins or 33
    mode = os.environ.get("QOB_DIST_EXCHANGE", DEFAULT_EXCHANGE)
    # slabs per rank: x, y (capi in direct mode: the exchange adds into the owners' y); + contributions (fused); + staging (nccl)
    # capi: contribution slabs (3 per rank) when they fit — the fold-in then pipelines behind the exchange and y needs no zero-fill:
    # 8 GPUs, N=33: 47.8 ms against 49.8 ms — and the direct mode (2 per rank) when they do not: N=33 on 2 GPUs
    direct = None
    if mode == "capi":
        want_direct = os.environ.get("QOB_DIST_DIRECT")
        if want_direct is not None:
            direct = want_direct != "0"
        else:
            direct = 3 * 16 * (1 << (n - p)) > 0.85 * free
    nbuf = {"capi": 2 if direct else 3, "fused": 3}.get(mode, 4)
    while nbuf * 16 * (1 << (n - p)) > 0.85 * free and n > 20:
        n -= 1
    nloc = n - p
    B, H = build_chain(Q, n)
    exchange = mode
    x = None
    if exchange == "capi":
        # the whole sharded apply behind the C ABI (qob_dist_*): the library owns planning, IPC mapping, barriers and streams
        from qob200.dist import DistLazySum

        try:
            sh = DistLazySum(H, rank, world, direct=True if direct else False)
            x = sh.x
        except Exception as e:     # CUDA IPC between the ranks' processes unavailable on this box
            if rank == 0:
                print(f"[bench] qob_dist_* unavailable ({type(e).__name__}: {e}); using the torch symmetric-memory orchestration", file=sys.stderr)
            exchange = "fused"
    if exchange != "capi":
        sh = ShardedLazySum(H, rank, world)
    if exchange == "fused":
        try:
            x = sh.empty_state()   # symmetric memory: peers load their tiles straight from this slab over NVLink
        except Exception as e:     # no symmetric-memory support on this box: NCCL all-to-all axis swaps instead
            if rank == 0:
                print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using the NCCL exchange", file=sys.stderr)
            exchange = "nccl"
    if x is None:
        x = torch.empty(1 << nloc, dtype=torch.complex128, device="cuda")
    Q.fill_state(x, 7, 2.0 ** (-n / 2), offset=rank << nloc)
    y = sh.y if exchange == "capi" and sh.direct else torch.empty(1 << nloc, dtype=torch.complex128, device="cuda")
    if exchange == "capi":
        exchange = "capi-direct" if sh.direct else "capi"
    plan = f"exchange={exchange}: " + sh.describe()

    def step():
        if exchange.startswith("capi"):
            sh.mul_(y, alpha, 0.0)
        elif exchange == "fused":
            sh.mul_fused_(y, x, alpha, 0.0)
        else:
            sh.mul_(y, x, alpha, 0.0)

    extra = {}
    if rank == 0 and free > 5.5 * 16 * (1 << nloc):
        # "1 GPU scaled": the same chain at the per-GPU slab size (N = nloc spins) on ONE GPU, timed in this run
        B1, H1 = build_chain(Q, nloc)
        x1, y1 = Q.Ket(B1), Q.Ket(B1)
        Q.fill_state(x1.data, 7, 2.0 ** (-nloc / 2))
        for _ in range(2):
            Q.mul_(y1, H1, x1, alpha, 0.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            Q.mul_(y1, H1, x1, alpha, 0.0)
        e1.record()
        torch.cuda.synchronize()
        ms1 = e0.elapsed_time(e1) / 3
        extra["single_gpu_same_slab"] = {"spins": nloc, "ms_per_step": ms1, "value": (1 << nloc) / (ms1 * 1e-3), "unit": UNIT,
                                         "note": "one GPU working on a chain of the per-GPU slab size; weak-scaling efficiency = value / (n_gpus * this)"}
        del x1, y1, H1
        torch.cuda.empty_cache()
    dist.barrier()
    run_common(args, Q, rank, world, step, x, y, n, nloc, plan, warmup, args.steps, peak, peak_src, H, alpha, sharded=sh, extra=extra)


PAULI = None


def spot_parity(ys, n, nloc, rank, alpha, seed=7, nsamples=64):
    """Exact per-amplitude oracle at sampled indices of this rank's slab: every output amplitude of the
    Heisenberg LazySum depends on <= 1 + n_terms inputs, and the input is the counter-based state both sides generate
    bit-identically (qob_fill_state / orc_state_at).  Returns the max relative error over the samples."""
    global PAULI
    import numpy as np
    import torch

    from oracle import qob_oracle as O

    if PAULI is None:
        PAULI = [np.array([[0, 1], [1, 0]], dtype=complex), np.array([[0, -1j], [1j, 0]], dtype=complex),
                 np.array([[1, 0], [0, -1]], dtype=complex)]
    spec = chain_spec(n)
    scale = 2.0 ** (-n / 2)
    D = 1 << nloc
    rng = np.random.default_rng(4242 + rank)
    local = [0, 1, 7, 8, 4095, 4096, D - 1, D - 2, D // 2, D // 2 - 1, (1 << (nloc - 1)) + 5, 0x5555555 % D, 0xAAAAAAA % D]
    local += [int(v) for v in rng.integers(0, D, max(0, nsamples - len(local)))]
    got = ys[torch.tensor(local, device=ys.device)].cpu().numpy()
    base = rank << nloc
    worst, ref_max = 0.0, 0.0
    refs = []
    for li in local:
        index = base + li
        acc = 0.0 + 0.0j
        for c, idx, a in spec:
            A = PAULI[a]
            k1, k2 = idx[0] - 1, idx[1] - 1
            i1, i2 = (index >> k1) & 1, (index >> k2) & 1
            for j1 in (0, 1):
                for j2 in (0, 1):
                    w = A[i1, j1] * A[i2, j2]
                    if w != 0:
                        jidx = (index & ~((1 << k1) | (1 << k2))) | (j1 << k1) | (j2 << k2)
                        acc += c * w * O.state_at(seed, jidx, scale)
        refs.append(alpha * acc)
    refs = np.array(refs)
    ref_max = float(np.abs(refs).max())
    worst = float(np.abs(got - refs).max() / ref_max) if ref_max > 0 else float(np.abs(got).max())
    return worst, len(local)


def extra_configs(Q, peak):
    """BASELINE configs 1-3 (and config 2 in its bandwidth regime) measured in the same run, each with its oracle check and the
    roofline that bounds it: launch latency for the tiny states, HBM for the large sparse gemm!, FP64 tensor
    (DMMA) for the dense d=48 factors."""
    import bench_configs as BC
    from oracle import qob_oracle as O

    recs = []
    try:
        BC.config1(Q, O, recs.append)
        BC.config2(Q, O, recs.append, cutoffs=(64, 4096))
        BC.config3(Q, O, recs.append)
    except Exception as e:  # noqa: BLE001
        recs.append({"config": "error", "error": f"{type(e).__name__}: {e}"})
    out = []
    for r in recs:
        c = r.get("config", "")
        e = {"config": c, "rel_err_vs_oracle": r.get("rel_err_vs_oracle"), "speedup_vs_cpu_port": r.get("speedup_vs_cpu_port")}
        if c.startswith("1:"):
            e.update(us_per_mul=r["us_per_mul"], us_per_mul_cuda_graph_replay=r.get("us_per_mul_cuda_graph_replay"),
                     amplitude_updates_per_s=r["amplitude_updates_per_s"],
                     roofline={"bound": "launch latency", "note": "64 KiB state, one fused launch: a bandwidth fraction is meaningless"})
        elif c.startswith("2:"):
            e.update(us_per_commutator=r["us_per_commutator"], us_per_commutator_cuda_graph_replay=r.get("us_per_commutator_cuda_graph_replay"),
                     roofline=({"bound": "hbm", "achieved": r["GBps"], "peak": peak, "unit": "GB/s", "frac": r["GBps"] / peak,
                                "algorithmic_GB": r["algorithmic_GB"]} if r.get("bound") == "HBM" else
                               {"bound": "launch latency", "note": "264 KiB operands: a bandwidth fraction is meaningless"}))
        elif c.startswith("3:"):
            e.update(ms_per_mul=r["ms_per_mul"], fp64_TFLOPs=r["fp64_TFLOPs"], plan=r.get("plan"),
                     roofline={"bound": "tensor", "achieved": r["fp64_TFLOPs"], "peak": 37.2, "unit": "TFLOP/s",
                               "frac": r["fp64_TFLOPs"] / 37.2,
                               "peak_source": "register-only DMMA.8x8x4 loop measured on this pool's B200 (tools/fp64_peak.cu); "
                                              "tcgen05 has no FP64 kind"})
        else:
            e.update(r)
        out.append(e)
    return out


def run_common(args, Q, rank, world, step, xs, ys, n, nloc, plan, warmup, steps, peak, peak_src, H, alpha, sharded=None, extra=None):
    import numpy as np
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    nterms = len(H.operators)
    for _ in range(warmup):
        step()
    barrier()
    # ---- timed region: exactly `steps` steps
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    if sharded is not None and hasattr(sharded, "exchange_stats"):
        sharded.time_exchange = True
        sharded._ex_events = []
    l0 = Q.launch_count()
    Q.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    nvlink = None
    if sharded is not None and hasattr(sharded, "exchange_stats"):
        sharded.time_exchange = False
        ex_ms, per_dir = sharded.exchange_stats()
        if ex_ms:
            t = torch.tensor([ex_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ex_ms = float(t.item())
            gbps = per_dir / 1e9 / (ex_ms * 1e-3)
            nvlink = {"bytes_per_direction_per_gpu": per_dir, "exchange_ms": ex_ms, "GBps_per_direction": gbps,
                      "frac_of_measured_770": gbps / 770.0, "frac_of_nominal_900": gbps / 900.0,
                      "note": "fused exchange kernel(s) of one mul!, CUDA events on their stream, max over ranks; runs beside the "
                              "local tile passes; 770 GB/s = measured peer-copy rate per direction (B200_PROFILING.md)"}
    Q.profile_enable(False)
    prof = Q.profile_read()
    launches = Q.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / steps
    amps = float(1 << n)
    value = amps / (ms_step * 1e-3)

    # ---- end-to-end through the public host-buffer API: pinned host -> device, mul!, device -> pinned host
    # N=1: a batch of K host kets per call (the columns of a host-resident D x K matrix, mul!(Y, H, X)): libqob200 streams
    # the kets through the device, so the upload of ket j+1, the kernels of ket j and the download of ket j-1 overlap and
    # the call is bound by one direction of PCIe.  The single-ket call (upload, kernels, download back to back) is timed
    # too and reported as e2e.single_ket_ms.
    e2e_steps = max(1, min(3, steps)) if world == 1 else 1
    slab_bytes = 16 * (1 << nloc)
    e2e_s, e2e_err, t_e2e, e2e_single_s, kets = None, None, [], None, 1
    if rank == 0:
        print(f"[bench] timed region: {ms_step:.3f} ms per step; end-to-end leg next", file=sys.stderr, flush=True)
    try:
        if os.environ.get("QOB_BENCH_SKIP_E2E"):   # tuning sweeps only; the contract run never sets it
            raise RuntimeError("skipped (QOB_BENCH_SKIP_E2E)")
        if sharded is None:
            kets = max(1, int(os.environ.get("QOB_BENCH_E2E_KETS", "4")))
            hx = torch.empty((kets, 1 << nloc), dtype=torch.complex128, pin_memory=True)
            hy = torch.empty((kets, 1 << nloc), dtype=torch.complex128, pin_memory=True)
            for k in range(kets):
                hx[k].copy_(xs)
            torch.cuda.synchronize()
            ts = []
            for i in range(3):
                t0 = time.perf_counter()
                Q.apply_host(H, hx[0].numpy(), alpha=alpha, beta=0.0, y=hy[0].numpy())
                ts.append(time.perf_counter() - t0)
            e2e_single_s = min(ts[1:])
            for i in range(e2e_steps + 1):
                t0 = time.perf_counter()
                Q.apply_host(H, hx.numpy().reshape(-1), alpha=alpha, beta=0.0, y=hy.numpy().reshape(-1), batch=kets)
                if i > 0:
                    t_e2e.append((time.perf_counter() - t0) / kets)
        else:
            # one pinned slab per rank serves both directions (upload, mul!, download are ordered on the stream); skip the leg
            # rather than let the kernel's OOM killer end the run when the ranks of this node cannot pin that much
            need = slab_bytes * world
            if need > 0.6 * host_memory_available():
                raise MemoryError(f"e2e leg needs {need >> 30} GiB of pinned host memory on this node")
            hx = torch.empty(1 << nloc, dtype=torch.complex128, pin_memory=True)
            hy = hx
            for i in range(e2e_steps + 1):
                hx.copy_(xs)               # (untimed) the shared pinned slab holds x again: the last download put y there
                torch.cuda.synchronize()
                barrier()
                t0 = time.perf_counter()
                xs.copy_(hx, non_blocking=True)
                step()
                hy.copy_(ys, non_blocking=True)
                barrier()
                if i > 0:
                    t_e2e.append(time.perf_counter() - t0)
        e2e_s = sum(t_e2e) / len(t_e2e)
        del hx, hy
    except Exception as e:  # e.g. the host cannot pin that much memory
        e2e_err = f"{type(e).__name__}: {e}"
    if world > 1:
        t = torch.tensor([e2e_s if e2e_s is not None else -1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mn = torch.tensor([e2e_s if e2e_s is not None else -1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        e2e_s = float(t.item()) if float(mn.item()) > 0 else None

    # ---- parity where the driver sees it: exact per-amplitude oracle at 64 sampled indices of EVERY rank's slab (the last
    # step left y = alpha * H x in ys); the run fails above 1e-12
    step()
    torch.cuda.synchronize()
    par_err, par_n = spot_parity(ys, n, nloc, rank if sharded is not None else 0, alpha)
    if world > 1:
        t = torch.tensor([par_err], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        par_err = float(t.item())
        par_n *= world
    if rank != 0:
        return
    # ---- roofline: achieved = algorithmic bytes of ONE complete mul! (32 B per amplitude: x read once, y
    # written once) / time of the mul!, against the measured HBM peak.  The per-launch figures (live CUDA events around every
    # launch of the tile kernel) use each launch's DRAM-level algorithmic bytes: 32 B/amplitude for the launch that defines y,
    # 48 B/amplitude for a launch that accumulates into it — however many tile passes are chained through L2 inside it.
    tot_ms = sum(p[0] for p in prof)
    by_pass = {}
    for ms, pi, by in prof:
        by_pass.setdefault(pi, []).append((ms, by))
    per_launch = [{"launch": k, "ms": sum(m for m, _ in v) / len(v), "algorithmic_GB": v[0][1] / 1e9,
                   "GBps": (v[0][1] / 1e9) / (sum(m for m, _ in v) / len(v) * 1e-3),
                   "frac": (v[0][1] / 1e9) / (sum(m for m, _ in v) / len(v) * 1e-3) / peak}
                  for k, v in sorted(by_pass.items())]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if world == 1 and n == 28 and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_mul")
        except Exception:
            traffic = None
    achieved = 32.0 * (1 << nloc) / 1e9 / (ms_step * 1e-3)
    kernel = "qreg_kernel" if "qreg[" in plan else "qtile_kernel"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "definition": "32 B per amplitude (x read once, y written once; beta = 0) x 2^n amplitudes per GPU / ms_per_step "
                              "(SURVEY.md section 8d); traffic = ncu dram bytes of one whole mul! (all launches)",
                "kernel": f"{kernel} (all launches of the fused LazySum apply; {len(per_launch)} per mul!)",
                "peak_source": peak_src, "algorithmic_bytes_per_mul": 32.0 * (1 << nloc),
                "kernel_share_of_step": tot_ms / ms_total if ms_total > 0 else None, "per_launch": per_launch}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Heisenberg XYZ spin-1/2 chain N={n} periodic, LazySum of {nterms} LazyTensor terms "
                               f"(sparse sigma factors), mul!(y,H,x,alpha,0) on Ket of 2^{n} ComplexF64",
                   "state_bytes": 16 * (1 << n), "l2": "operands (>= 4 GiB per GPU) exceed the 126 MB L2; no flush needed",
                   "plan": plan, "parallelism": "single GPU" if world == 1 else
                   f"state sharded on the top {world.bit_length() - 1} spin axes ({16 * (1 << nloc) / 2**30:.0f} GiB slab per GPU, the largest "
                   f"chain that fits); remote terms " + (
                       "exchanged by the fused peer-memory tile pass over NVLink, orchestrated behind the C ABI (qob_dist_*); the pass "
                       "adds into the owners' result slabs (2 slabs per rank)" if plan.startswith("exchange=capi-direct") else
                       "exchanged by the fused peer-memory tile pass over NVLink, orchestrated behind the C ABI (qob_dist_*)"
                       if plan.startswith("exchange=capi") else
                       "exchanged by the fused peer-memory tile pass over NVLink (torch symmetric memory orchestration)"
                       if plan.startswith("exchange=fused") else "through NCCL all-to-all axis swaps")},
        "term_updates_per_s": value * nterms,
        "hbm_GBps_algorithmic": 32.0 * (1 << nloc) * world / 1e9 / (ms_step * 1e-3),
        "clocks": clocks,
        "e2e": {"value": (amps / e2e_s) if e2e_s else None, "unit": UNIT, "h2d_bytes_per_step": slab_bytes * world,
                "d2h_bytes_per_step": slab_bytes * world, "ms_per_step": (1e3 * e2e_s) if e2e_s else None,
                "steps": len(t_e2e) * kets,
                **({"mode": f"qob_op_apply_host on a batch of {kets} pinned host kets per call, upload / kernels / download "
                            f"pipelined across kets inside the library; every ket is uploaded and its result downloaded in "
                            f"the timed region", "single_ket_ms": 1e3 * e2e_single_s} if sharded is None and e2e_single_s else {}),
                **({"error": e2e_err} if e2e_err else {})},
        "gpu_launches": launches,
        "roofline": roofline,
        "parity": {"max_rel_err": par_err, "n_samples": par_n, "tolerance": 1e-12,
                   "oracle": "exact per-amplitude evaluation of the LazySum definition on the counter-based input "
                             "(oracle/qob_oracle.c: orc_state_at), slab / tile boundaries + random indices on every rank"},
    }
    if nvlink is not None:
        line["nvlink"] = nvlink
    if extra:
        line.update(extra)
        if "single_gpu_same_slab" in extra:
            line["weak_scaling_efficiency_vs_same_slab"] = value / (world * extra["single_gpu_same_slab"]["value"])
    if world == 1 and not args.no_extra_configs:
        del xs, ys
        torch.cuda.empty_cache()
        line["extra_configs"] = extra_configs(Q, peak)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(n, nterms)
    else:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": "reported at N=1 only"}
    print(json.dumps(line), flush=True)
    if not (par_err <= 1e-12):
        print(f"[bench] PARITY FAILURE: max relative error {par_err:.3e} at {par_n} sampled amplitudes", file=sys.stderr)
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qob200", choices=["qob200", "reference"])
    ap.add_argument("--spins", type=int, default=0, help="override the chain length (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip BASELINE configs 1-3 (N=1 only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world if world > 1 else args.gpus)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        try:
            run_product_dist(args, rank, world, local_rank)
        finally:
            dist.destroy_process_group()
    else:
        if args.gpus > 1:
            print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun (WORLD_SIZE={world})"}))
            sys.exit(2)
        run_product(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
