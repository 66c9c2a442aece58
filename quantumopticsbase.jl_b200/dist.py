"""Sharded (multi-GPU) LazySum apply: one process per GPU, torch.distributed for the plumbing.

The reference has no distributed path at all: a state must fit one array.  Here a
spin-1/2 state of N subsystems is sharded on its highest-stride axes: rank r of P = 2^p holds the
contiguous slab [r*2^(N-p), (r+1)*2^(N-p)) of the reference's column-major array, i.e. index bits
N-p..N-1 are the rank.

  * Terms whose OFF-DIAGONAL factors all sit on local bits run with no communication; diagonal factors
    on sharded bits (sigma_z, number) only select a rank-dependent weight (`hi_value` in the tile plan).
  * The remaining terms (sigma_x/sigma_y on a sharded axis) run in a SWAPPED layout: the p sharded bits
    trade places with a window of p local bits that those terms do not touch, by ONE all-to-all each
    way (`axis_swap`, NCCL all_to_all_single over NVLink).  The swap is an involution.

Compute kernels are libqob200's tile programs (`qob_layout_plan_*`); this module only decides layouts
and moves data.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib
from ._lib import c64, lib
from .operators import LazySum, handle


def swap_window(nloc: int, p: int, touched_mask: int) -> int:
    """Lowest bit of the highest window of p local bits that no swapped term touches."""
    for c in range(nloc - p, -1, -1):
        win = ((1 << p) - 1) << c
        if not (touched_mask & win):
            return c
    raise _lib.MethodError(f"no free window of {p} local bits for the axis swap")


def swapped_bitpos(n: int, nloc: int, p: int, s: int):
    """Virtual position of every subsystem bit after the axis swap: sharded bits move to [s, s+p),
    the window bits become the rank bits."""
    pos = list(range(n))
    for k in range(nloc, n):
        pos[k] = s + (k - nloc)
    for k in range(s, s + p):
        pos[k] = nloc + (k - s)
    return pos


def axis_swap(t, s: int, p: int, out, group=None, async_op=False):
    """All-to-all axis swap of a local slab `t` (2^nloc complex128): local bits [s, s+p) <-> rank bits.

    Viewed as (hi, P, lo) with lo = 2^s, block [h, d, :] goes to rank d and lands at [h, src, :].
    Works on CUDA tensors (NCCL) and on CPU tensors (gloo, used by the world_size-2 tests).
    async_op=True returns the list of work handles (the collectives run on NCCL's stream, ordered after the
    work already queued on the current stream); call `.wait()` on each before consuming `out`."""
    import torch
    import torch.distributed as dist

    P = 1 << p
    lo = 1 << s
    n = t.numel()
    hi = n // (P * lo)
    tin = torch.view_as_real(t).view(hi, P * lo * 2)
    tout = torch.view_as_real(out).view(hi, P * lo * 2)
    if P == 1:
        tout.copy_(tin)
        return [] if async_op else out
    works = []
    for h in range(hi):
        w = dist.all_to_all_single(tout[h], tin[h], group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works if async_op else out


class ShardedLazySum:
    """`mul!(y, H, x, alpha, beta)` for a Ket sharded over the ranks of a torch.distributed group.

    Schedule of one apply (the two all-to-alls overlap the local tile passes):
        comm:     x' = swap(x) ...................           y'' = swap(y') ...................
        compute:  y = alpha*H_A x + beta*y            y' = alpha*H_remote x'   y += alpha*H_B x        y += y''
    H_A / H_B: the communication-free terms, split at an index bit so that both halves cover a swap."""

    def __init__(self, H: LazySum, rank: int, world: int, group=None, ctx=None, overlap=True, swap_sms=None):
        assert world & (world - 1) == 0, "world size must be a power of two"
        self.H, self.rank, self.world, self.group = H, rank, world, group
        self.n = len(H.basis_l.shape)
        self.p = world.bit_length() - 1
        self.nloc = self.n - self.p
        self.ctx = ctx
        self.overlap = overlap
        self.h = handle(H, ctx)
        nterms = len(H.operators)
        od, al = C.c_uint64(), C.c_uint64()
        lowmask = (1 << self.nloc) - 1
        # off-diagonal terms entirely below split_bit -> group A (runs beside the exchange); the rest (one window pass
        # over the top local bits) -> group B, which also folds the received contributions in
        split_bit = int(os.environ.get("QOB_DIST_SPLIT_BIT", self.nloc - 6))
        sel = {k: (C.c_uint8 * max(nterms, 1))() for k in ("A", "B", "R")}
        touched = 0
        self.n_local = self.n_remote = 0
        counts = {"A": 0, "B": 0, "R": 0}
        for i in range(nterms):
            _lib.check(lib.qob_lazysum_term_masks(self.h, i, C.byref(od), C.byref(al)))
            if od.value & ~lowmask:
                k = "R"
                touched |= al.value
            elif od.value == 0 or (od.value >> split_bit) == 0:
                k = "A"          # diagonal terms cost no pass of their own: keep them with the first group
            else:
                k = "B"
            sel[k][i] = 1
            counts[k] += 1
        self.n_remote = counts["R"]
        self.n_local = counts["A"] + counts["B"]
        if not self.n_remote or not overlap:   # nothing to overlap with: one local plan
            for i in range(nterms):
                if sel["B"][i]:
                    sel["A"][i], sel["B"][i] = 1, 0
            counts["A"] += counts["B"]
            counts["B"] = 0
        self.plan_info = {}
        ident = list(range(self.n))

        def make(select, bitpos):
            pid = C.c_int32()
            pos = (C.c_int32 * self.n)(*bitpos)
            _lib.check(lib.qob_layout_plan_create(self.h, self.nloc, pos, C.c_uint64(rank), select, C.byref(pid)))
            self.plan_info[pid.value] = dict(bitpos=list(bitpos), hi_value=rank, select=[bool(v) for v in select][:nterms])
            return pid.value

        self.plan_local = make(sel["A"], ident)          # always exists: it also carries the beta update
        self.plan_local_b = make(sel["B"], ident) if counts["B"] else None
        self.plan_swapped = None
        self.swap_lo = None
        if self.n_remote:
            self.swap_lo = swap_window(self.nloc, self.p, touched & lowmask)
            self.plan_swapped = make(sel["R"], swapped_bitpos(self.n, self.nloc, self.p, self.swap_lo))
        self._buf = None
        self.nchunks, self.chunk_mask = 1, 0
        if os.environ.get("QOB_DIST_CHUNKS", "4") not in ("0", "1"):
            self._setup_chunks(int(os.environ.get("QOB_DIST_CHUNKS", "4")))
        # fused exchange (symmetric memory): peers' pointer tables, keyed by the local tensor's data_ptr
        self._symm = {}
        self._zbuf = None
        self.swap_sms = int(os.environ.get("QOB_DIST_SWAP_SMS", "32")) if swap_sms is None else int(swap_sms)
        self.time_exchange = False     # bench.py: bracket the exchange kernels with CUDA events on their stream
        self._ex_events = []
        self.local_budget = int(os.environ.get("QOB_DIST_LOCAL_SMS", "0"))   # 0: full grid (one CTA per tile)

    # ------------------------------------------------------------------ fused exchange over NVLink peer memory
    def empty_state(self, device=None):
        """A slab of 2^nloc ComplexF64 in SYMMETRIC memory (torch.distributed._symmetric_memory): every rank can address
        every other rank's slab, which is what lets the remote-term pass load its tiles straight from the owners' HBM.
        Collective: all ranks must call it in the same order."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        raw = symm_mem.empty((1 << self.nloc) * 2, dtype=torch.float64, device=dev)
        hdl = symm_mem.rendezvous(raw, self.group if self.group is not None else dist.group.WORLD)
        t = torch.view_as_complex(raw.view(1 << self.nloc, 2))
        self._symm[t.data_ptr()] = (hdl, [int(p) for p in hdl.buffer_ptrs], raw)
        return t

    def _peer_table(self, ptrs):
        return (C.c_void_p * len(ptrs))(*ptrs)

    def _apply_ex(self, plan, alpha, x, beta, y, zadd=None, peers=None, sm_budget=0, chunk=(0, 1)):
        import torch

        handle(self.H, self.ctx)
        npeers, xp, yp, shift = 0, None, None, 0
        if peers is not None:
            xp, yp = self._peer_table(peers[0]), self._peer_table(peers[1])
            npeers, shift = len(peers[0]), self.swap_lo
        _lib.check(lib.qob_layout_plan_apply_ex(
            self.h, plan, c64.of(alpha), C.c_void_p(x.data_ptr() if x is not None else 0), c64.of(beta),
            C.c_void_p(y.data_ptr() if y is not None else 0), C.c_void_p(zadd.data_ptr() if zadd is not None else 0),
            npeers, xp, yp, shift, int(sm_budget), int(chunk[0]), int(chunk[1]),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def _plan_info(self, plan):
        n, m = C.c_int32(), C.c_uint64()
        _lib.check(lib.qob_layout_plan_info(self.h, plan, C.byref(n), C.byref(m)))
        return n.value, m.value

    def _setup_chunks(self, want=4):
        """Pipeline the fold-in behind the exchange: when the remote-term plan and the last local group are single-pass
        plans, both are chunked on the same (top common fixed) index bits, so chunk c of the local group can run as soon
        as chunk c of the exchange has landed on every rank."""
        self.nchunks = 1
        if self.plan_swapped is None or self.plan_local_b is None or not self.overlap:
            return
        (ns, fs), (nb, fb) = self._plan_info(self.plan_swapped), self._plan_info(self.plan_local_b)
        if ns != 1 or nb != 1:
            return
        window = ((1 << self.p) - 1) << self.swap_lo
        common = fs & fb & ((1 << self.nloc) - 1) & ~window
        bits = [b for b in range(self.nloc - 1, -1, -1) if common >> b & 1][: max(1, want.bit_length() - 1)]
        if len(bits) < 1:
            return
        mask = 0
        for b in bits:
            mask |= 1 << b
        _lib.check(lib.qob_layout_plan_set_chunk_bits(self.h, self.plan_swapped, C.c_uint64(mask)))
        _lib.check(lib.qob_layout_plan_set_chunk_bits(self.h, self.plan_local_b, C.c_uint64(mask)))
        self.nchunks = 1 << len(bits)
        self.chunk_mask = mask

    def mul_fused_(self, y, x, alpha=1.0, beta=0.0):
        """Same result as mul_, with the exchange fused into the compute kernel: the remote-term pass reads x tiles from
        the peers' slabs and writes its results into the owners' contribution buffers in ONE kernel over NVLink
        (no staging copy, no NCCL), running beside the first group of local passes on its own share of the SMs; the last
        local pass folds the received contributions in (zadd).  `x` must come from `empty_state()`."""
        import torch

        alpha, beta = complex(alpha), complex(beta)
        if self.plan_swapped is None or alpha == 0:
            return self.mul_(y, x, alpha, beta)
        if x.data_ptr() not in self._symm:
            raise _lib.ArgumentError("mul_fused_ needs a state allocated with empty_state() (symmetric memory)")
        if self._zbuf is None:
            self._zbuf = self.empty_state()
            self._side = torch.cuda.Stream(priority=-1)   # the exchange kernel's CTAs get free SM slots first
        zh, zptrs, _ = self._symm[self._zbuf.data_ptr()]
        _, xptrs, _ = self._symm[x.data_ptr()]
        main = torch.cuda.current_stream()
        total = torch.cuda.get_device_properties(x.device).multi_processor_count
        k = max(4, min(self.swap_sms, total // 2))
        two_groups = self.plan_local_b is not None
        if self.overlap and two_groups:
            side = self._side
            nc = self.nchunks
            side.wait_stream(main)                     # x is ready on this rank
            events = []
            with torch.cuda.stream(side):
                zh.barrier(channel=0)                  # ... and on every rank; last call's contributions are consumed
                if self.time_exchange:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record(side)
                for c in range(nc):
                    self._apply_ex(self.plan_swapped, alpha, None, 0.0, None, peers=(xptrs, zptrs), sm_budget=k, chunk=(c, nc))
                    zh.barrier(channel=1)              # chunk c of every rank's contributions has landed
                    ev = torch.cuda.Event()
                    ev.record(side)
                    events.append(ev)
                if self.time_exchange:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record(side)
                    self._ex_events.append((e0, e1))
            # beside the exchange: the local passes use a full grid; the exchange kernel is persistent with k*occupancy
            # CTAs on a high-priority stream, so it keeps its share of the slots while local CTAs come and go
            # sm_budget = -k: "all but k SMs": the exchange pass (round-2 kernel, one persistent CTA per SM on k SMs) and the
            # local passes (the same kernel on the other SMs) never share an SM, so the exchange keeps its NVLink rate
            self._apply_ex(self.plan_local, alpha, x, beta, y, sm_budget=self.local_budget if self.local_budget else -k)
            for c in range(nc):                        # fold the contributions in, chunk by chunk, behind the exchange
                main.wait_event(events[c])
                self._apply_ex(self.plan_local_b, alpha, x, 1.0, y, zadd=self._zbuf, chunk=(c, nc))
        else:
            zh.barrier(channel=0)
            self._apply_ex(self.plan_swapped, alpha, None, 0.0, None, peers=(xptrs, zptrs))
            zh.barrier(channel=1)
            if two_groups:
                self._apply_ex(self.plan_local, alpha, x, beta, y)
                self._apply_ex(self.plan_local_b, alpha, x, 1.0, y, zadd=self._zbuf)
            else:
                self._apply_ex(self.plan_local, alpha, x, beta, y, zadd=self._zbuf)
        return y

    def exchange_stats(self):
        """(mean ms of the timed exchanges, bytes that crossed NVLink per direction on this GPU per apply): the remote-term pass
        loads the (1 - 1/P) share of its swapped-layout x tiles from the peers and stores the same share of its results into
        the peers' contribution buffers; the peers do the same to this GPU."""
        import torch

        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in self._ex_events]
        self._ex_events = []
        per_dir = 2.0 * (1.0 - 1.0 / self.world) * 16.0 * (1 << self.nloc)
        return (sum(ms) / len(ms) if ms else None), per_dir

    def describe(self):
        buf = C.create_string_buffer(1 << 14)

        def d(pid):
            _lib.check(lib.qob_layout_plan_describe(self.h, pid, buf, len(buf)))
            return buf.value.decode()
        out = f"local[{self.n_local} terms]: {d(self.plan_local)}"
        if self.plan_local_b is not None:
            out += f" + {d(self.plan_local_b)}"
        if self.plan_swapped is not None:
            out += f" | swapped[{self.n_remote} terms, window bit {self.swap_lo}]: {d(self.plan_swapped)}"
        return out

    def _apply(self, plan, alpha, x, beta, y):
        import torch

        # coefficients may have been mutated (TimeDependentSum): handle() re-sends them
        handle(self.H, self.ctx)
        _lib.check(lib.qob_layout_plan_apply(self.h, plan, c64.of(alpha), C.c_void_p(x.data_ptr()), c64.of(beta),
                                             C.c_void_p(y.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def mul_(self, y, x, alpha=1.0, beta=0.0):
        """y_local = alpha * (H x)_local + beta * y_local;  x, y: torch complex128 slabs of 2^nloc amplitudes."""
        import torch

        assert x.numel() == 1 << self.nloc and y.numel() == 1 << self.nloc
        alpha, beta = complex(alpha), complex(beta)
        if self.plan_swapped is None or alpha == 0:
            self._apply(self.plan_local, alpha, x, beta, y)
            if self.plan_local_b is not None and alpha != 0:
                self._apply(self.plan_local_b, alpha, x, 1.0, y)
            return y
        if self._buf is None or self._buf[0].numel() != x.numel() or self._buf[0].device != x.device:
            self._buf = (torch.empty_like(x), torch.empty_like(x))
        b1, b2 = self._buf
        use_async = self.overlap and x.is_cuda
        works = axis_swap(x, self.swap_lo, self.p, b1, self.group, async_op=use_async)     # x in the swapped layout
        self._apply(self.plan_local, alpha, x, beta, y)                                     # overlaps the swap
        for w in (works if use_async else []):
            w.wait()
        self._apply(self.plan_swapped, alpha, b1, 0.0, b2)                                  # partial result, swapped layout
        works = axis_swap(b2, self.swap_lo, self.p, b1, self.group, async_op=use_async)     # back to the slab layout
        if self.plan_local_b is not None:
            self._apply(self.plan_local_b, alpha, x, 1.0, y)                                # overlaps the swap back
        for w in (works if use_async else []):
            w.wait()
        y.add_(b1)
        return y


class _DevBuf:
    """A raw device allocation seen through __cuda_array_interface__, so that torch can view it without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DistLazySum:
    """The sharded apply THROUGH THE C ABI (`qob_dist_*`, csrc/qob_dist.cu): planning, CUDA-IPC mapping of the peers' slabs,
    device-side barriers and the stream choreography of the fused exchange all live in libqob200.  This class only
    allocates, moves the 64-byte IPC handles between the processes (torch.distributed all_gather — any channel would do) and
    forwards `mul_`.  It is what a Julia / C host does with MPI or sockets instead.

        sh = DistLazySum(H, rank, world)     # collective
        sh.x                                 # this rank's slab of the state (torch complex128 view of library memory)
        sh.mul_(y, alpha, beta)              # y_local = alpha * (H x)_local + beta * y_local, collective
    """

    def __init__(self, H: LazySum, rank: int, world: int, group=None, ctx=None, direct=None):
        """direct: None = use the direct mode when the plan allows it (the result slab `self.y` is library memory mapped by the
        peers, the exchange adds into it, no contribution slab: 2 slabs per rank); False = contribution slabs (3 per rank,
        any result buffer)."""
        import torch
        import torch.distributed as dist

        self.H, self.rank, self.world, self.group = H, rank, world, group
        self.ctx = _lib.context() if ctx is None else ctx
        self.h = handle(H, self.ctx)
        d = C.c_void_p()
        _lib.check(lib.qob_dist_create(self.h, rank, world, C.byref(d)))
        self.d = d
        nloc, nrem, nch = C.c_int32(), C.c_int32(), C.c_int32()
        slab, flagb = C.c_int64(), C.c_int64()
        _lib.check(lib.qob_dist_info(d, C.byref(nloc), C.byref(nrem), C.byref(nch), C.byref(slab), C.byref(flagb)))
        self.nloc, self.n_remote, self.nchunks = nloc.value, nrem.value, nch.value
        self.n = len(H.basis_l.shape)
        self._own, self._peers = [], []
        self._timing = False
        self._ex_events = []
        cap = C.c_int32()
        _lib.check(lib.qob_dist_direct_capable(d, C.byref(cap)))
        self.direct = bool(cap.value) and direct is not False and self.n_remote > 0 and world > 1
        if direct and not self.direct:
            raise _lib.ArgumentError("direct mode is not available for this plan")

        def alloc(nbytes):
            p = C.c_void_p()
            _lib.check(lib.qob_dist_alloc(self.ctx, nbytes, C.byref(p)))
            self._own.append(p)
            return p

        def view(ptr, nbytes):
            return torch.view_as_complex(torch.as_tensor(_DevBuf(ptr.value, nbytes), device="cuda").view(-1, 2))

        px = alloc(slab.value)
        self.x = view(px, slab.value)
        self.y = None
        tables = [[px.value] * world, None, None]
        if self.n_remote and world > 1:
            # second slab: the result (direct mode) or the contribution buffer
            p2, pf = alloc(slab.value), alloc(flagb.value)
            if self.direct:
                self.y = view(p2, slab.value)
            torch.as_tensor(_DevBuf(pf.value, flagb.value), device="cuda").zero_()
            torch.cuda.synchronize()
            mine = torch.empty(3 * 64, dtype=torch.uint8)
            for k, p in enumerate((px, p2, pf)):
                hb = (C.c_uint8 * 64)()
                _lib.check(lib.qob_ipc_export(p, hb))
                mine[64 * k:64 * (k + 1)] = torch.frombuffer(bytearray(hb), dtype=torch.uint8)
            backend = dist.get_backend(group)
            dev = "cuda" if backend == "nccl" else "cpu"
            allh = torch.empty(world * 3 * 64, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, mine.to(dev), group=group)
            allh = allh.cpu().view(world, 3, 64)
            own = (px.value, p2.value, pf.value)
            tables = [[], [], []]
            for q in range(world):
                for k in range(3):
                    if q == rank:
                        tables[k].append(own[k])
                        continue
                    hb = (C.c_uint8 * 64)(*allh[q, k].tolist())
                    pp = C.c_void_p()
                    _lib.check(lib.qob_ipc_open(self.ctx, hb, C.byref(pp)))
                    self._peers.append(pp)
                    tables[k].append(pp.value)
            dist.barrier(group=group)   # every pad is zeroed and every mapping exists before the first apply
        arr = [None if t is None else (C.c_void_p * world)(*t) for t in tables]
        if self.direct:
            _lib.check(lib.qob_dist_bind(d, arr[0], None, arr[2]))
            _lib.check(lib.qob_dist_bind_result(d, arr[1]))
        else:
            _lib.check(lib.qob_dist_bind(d, arr[0], arr[1], arr[2]))

    def describe(self):
        buf = C.create_string_buffer(1 << 15)
        _lib.check(lib.qob_dist_describe(self.d, buf, len(buf)))
        return buf.value.decode()

    @property
    def time_exchange(self):
        return self._timing

    @time_exchange.setter
    def time_exchange(self, on):
        self._timing = bool(on)
        _lib.check(lib.qob_dist_exchange_timing(self.d, int(bool(on))))

    def exchange_stats(self):
        """(mean ms of the exchanges timed since the last call, bytes over NVLink per direction per apply on this GPU)"""
        ms, cnt, nb = C.c_double(), C.c_int32(), C.c_int64()
        _lib.check(lib.qob_dist_exchange_ms(self.d, C.byref(ms), C.byref(cnt), C.byref(nb)))
        return (ms.value if cnt.value else None), float(nb.value)

    def mul_(self, y=None, alpha=1.0, beta=0.0):
        """y_local = alpha * (H x)_local + beta * y_local, collective.  Direct mode: y is `self.y` (the default)."""
        import torch

        if y is None:
            y = self.y
        handle(self.H, self.ctx)   # coefficients may have been mutated (TimeDependentSum): re-sent here
        _lib.check(lib.qob_dist_apply(self.d, c64.of(complex(alpha)), c64.of(complex(beta)), C.c_void_p(y.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return y

    def close(self):
        import torch

        if getattr(self, "d", None):
            torch.cuda.synchronize()
            lib.qob_dist_destroy(self.d)
            self.d = None
            self.x = None
            self.y = None
            for p in self._peers:
                lib.qob_ipc_close(self.ctx, p)
            for p in self._own:
                lib.qob_dist_free(self.ctx, p)
            self._peers, self._own = [], []
