"""Layout optimiser: negative-sampling UMAP forces + InfoNCE + Adam on the device.

Host-side orchestration of kernels K7/K8/K9.  Mirrors UMAPMixture._train
(/root/reference/impl/model.py:396-481) for modes "fit" and "transform".

Two sample streams:
  * "host"   -- parity mode.  Every random draw is made on torch's global CPU generator with
                the reference's shapes, dtypes and order (model.py:432,
                :444, :373, :383), then uploaded; with torch.manual_seed the engine consumes
                exactly the reference's stream.
  * "device" -- throughput mode.  Philox4x32-10 counter streams evaluated inside the kernels;
                no host involvement inside an epoch, statistically equivalent draws.
"""
from __future__ import annotations

import os

import torch

from . import dist as D
from . import native, profiler
from .graph import Graph
from .native import check, lib, ptr, stream

BETA1, BETA2, EPS = 0.9, 0.999, 1e-8          # torch.optim.Adam defaults (model.py:403)
INFONCE_NEG = 9                               # n_neg + 1 draws per anchor (model.py:364,383)
INFONCE_CHUNK = 1000                          # model.py:369
INFONCE_TAU = 0.5                             # model.py:364
PUSH_TAIL_MAX_BYTES = 32 << 20                # multi-GPU: tables up to this size use the push form of the epoch tail
AUTO_WINDOW_MB = 80                           # p + g bytes of the tail rows of one force-kernel window (see _window_rows)


def default_stream() -> str:
    return os.environ.get("MMUMAP_SAMPLE_STREAM", "device")


class _Modality:
    """Device state of one table being optimised."""

    def __init__(self, embed: torch.Tensor, graph: Graph, batch_size: int, ref: torch.Tensor | None, flat, offset: int,
                 index: int, seed: int, host_stream: bool):
        dev = torch.device("cuda")
        n, d = embed.shape
        # p/g/m/v of all modalities live in four flat buffers: one all-reduce and one Adam launch per epoch
        self.p, self.g, self.m, self.v = (f[offset:offset + n * d].view(n, d) for f in flat)
        self.p.copy_(embed.detach().to(dev, torch.float32))
        self.graph = graph
        self.ref = None if ref is None else ref.detach().to(dev, torch.float32).contiguous()
        self.count = self.p.shape[0]
        self.dim = self.p.shape[1]
        self.batch_size = batch_size
        self.n_batches = (self.count + batch_size - 1) // batch_size
        self.rep_count = self.ref.shape[0] if self.ref is not None else self.count
        # every modality draws from its own Philox key (the reference draws an independent rand / randint per
        # modality, model.py:432,444): same key + same counters would give edge position p the same uniforms and
        # the same negatives in every modality
        self.seed = (int(seed) + index * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        self.batch_kept = torch.zeros(self.n_batches, dtype=torch.int32, device=dev)
        # multi-GPU: this rank owns a range of row-batches, i.e. a contiguous range of edges, chosen so that the
        # ranks' expected kept-edge counts (sum of weights, model.py:432) are equal
        w, r = D.world(), D.rank()
        if w == 1:
            self.b_lo, self.b_hi = 0, self.n_batches
            self.e_lo, self.e_hi = 0, graph.nnz
        else:
            ends = torch.arange(1, self.n_batches + 1, device=graph.rowptr.device, dtype=torch.int64) * batch_size
            edge_end = graph.rowptr[ends.clamp(max=self.count)]
            csum = torch.cumsum(graph.val.double(), 0)
            cum_w = torch.where(edge_end > 0, csum[(edge_end - 1).clamp(min=0)], torch.zeros_like(edge_end, dtype=torch.float64))
            self.b_lo, self.b_hi = D.balanced_batch_range(cum_w.cpu().tolist(), r, w)
            lo_row = min(self.b_lo * batch_size, self.count)
            hi_row = min(self.b_hi * batch_size, self.count)
            self.e_lo, self.e_hi = (int(v) for v in graph.rowptr[[lo_row, hi_row]].tolist())
        n_edges = self.e_hi - self.e_lo
        # E[kept edges per epoch] = this rank's sum of weights (Bernoulli(w), model.py:432); the record list holds
        # that plus 10 standard deviations (variance <= sum w (1 - w) <= sum w), never more than the edge count
        self.expected_kept = float(graph.val[self.e_lo:self.e_hi].double().sum().item()) if n_edges else 0.0
        cap = n_edges if host_stream else min(n_edges, int(self.expected_kept + 10.0 * self.expected_kept ** 0.5) + 4096)
        self.capacity = max(cap, 1)
        self.kept_rec = torch.empty((self.capacity, 4), dtype=torch.int32, device=dev)
        self.kept_hdr = self._new_hdr()
        self.kept_pos = None                 # host stream: uploaded positions of the replayed draws
        self.window_rows = None              # tail-window size of the force kernel, decided at the first launch
        # host copies for the replayed stream
        self._w_cpu = None
        self._rowptr_cpu = None

    def _new_hdr(self):
        return torch.tensor([0, self.capacity, 0, 0], dtype=torch.int32, device="cuda")

    def host_arrays(self):
        if self._w_cpu is None:
            self._w_cpu = self.graph.val.cpu()
            self._rowptr_cpu = self.graph.rowptr.cpu()
        return self._w_cpu, self._rowptr_cpu

    def load_host_draws(self, kept: torch.Tensor, counts: torch.Tensor):
        """Host sample stream: upload the replayed kept positions / per-batch counts and build the records."""
        dev = self.p.device
        n = int(kept.numel())
        if self.kept_pos is None or self.kept_pos.numel() < max(n, 1):
            self.kept_pos = torch.empty(max(n, 1, self.capacity), dtype=torch.int32, device=dev)
        self.kept_pos[:n].copy_(kept.to(torch.int32).pin_memory(), non_blocking=True)
        self.batch_kept.copy_(counts.to(torch.int32).pin_memory(), non_blocking=True)
        g = self.graph
        check(lib().mmu_edge_records(ptr(g.row), ptr(g.col), ptr(self.kept_pos), n, self.batch_size, ptr(self.kept_rec),
                                     ptr(self.kept_hdr), stream()), "mmu_edge_records")
        return n


def replay_host_draws(mod: _Modality, num_rep: int):
    """Issue the reference's per-batch draws (model.py:423-444) on the CPU generator.
    Returns (kept_pos int32 [kept], neg int32 [kept, num_rep], batch_kept int32 [n_batches])."""
    w_cpu, rowptr = mod.host_arrays()
    kept_chunks, neg_chunks, counts = [], [], []
    for j in range(0, mod.count, mod.batch_size):
        end = min(j + mod.batch_size, mod.count)
        lo, hi = int(rowptr[j]), int(rowptr[end])
        keep = torch.rand(hi - lo) < w_cpu[lo:hi]                       # model.py:432
        pos = torch.nonzero(keep).flatten() + lo
        num_pairs = pos.numel()
        neg = torch.randint(0, mod.rep_count, (num_pairs, num_rep))     # model.py:444
        kept_chunks.append(pos)
        neg_chunks.append(neg)
        counts.append(num_pairs)
    counts_t = torch.tensor(counts, dtype=torch.int32)
    if D.world() > 1:
        # every rank replays the whole stream (same generator state everywhere) and keeps its batches
        keep = range(mod.b_lo, mod.b_hi)
        kept_chunks = [kept_chunks[b] for b in keep] or [torch.zeros(0, dtype=torch.int64)]
        neg_chunks = [neg_chunks[b] for b in keep] or [torch.zeros((0, num_rep), dtype=torch.int64)]
    kept = torch.cat(kept_chunks).to(torch.int32)
    neg = torch.cat(neg_chunks).to(torch.int32).reshape(-1, num_rep)
    return kept, neg, counts_t


def replay_infonce_draws(num: int):
    """model.py:373 (randperm) and :383 (randint per chunk of 1000 anchors)."""
    perm = torch.randperm(num)
    negs = [torch.randint(0, num, (min(s + INFONCE_CHUNK, num) - s, INFONCE_NEG))
            for s in range(0, num, INFONCE_CHUNK)]
    neg = torch.cat(negs) if negs else torch.zeros((0, INFONCE_NEG), dtype=torch.int64)
    return perm.to(torch.int32), neg.to(torch.int32)


_PEER_CACHE: dict = {}


class LayoutOptimizer:
    def __init__(self, embeds, graphs, a: float, b: float, num_rep: int, lr: float, alpha: float,
                 batch_size: int, mode: str = "fit", refs=None, sample_stream: str | None = None,
                 seed: int | None = None, track_loss: bool = False, sigmas=None, rhos=None, norm_batches=None):
        native.require_cuda()
        if mode not in ("fit", "transform", "invert"):
            raise ValueError(f"Invalid mode: {mode}")
        if mode == "invert" and (refs is None or sigmas is None or rhos is None):
            raise ValueError("invert mode needs the target data rows and their fit-time sigma / rho")
        self.sigmas = None if sigmas is None else [t.detach().to("cuda", torch.float32).contiguous() for t in sigmas]
        self.rhos = None if rhos is None else [t.detach().to("cuda", torch.float32).contiguous() for t in rhos]
        self.mode = mode
        self.a, self.b = float(a), float(b)
        self.num_rep, self.lr, self.alpha = int(num_rep), float(lr), float(alpha)
        self.sample_stream = sample_stream or default_stream()
        if self.sample_stream not in ("host", "device"):
            raise ValueError(f"unknown sample stream {self.sample_stream!r}")
        dev = torch.device("cuda")
        # every table starts on a 256-byte boundary of the flat buffers (vector loads / red.v4)
        sizes = [-(-(int(e.shape[0]) * int(e.shape[1])) // 64) * 64 for e in embeds]
        total = sum(sizes)
        self.flat = tuple(torch.zeros(max(total, 1), dtype=torch.float32, device=dev) for _ in range(4))   # p, g, m, v
        self.total = total
        self.peer = self._peer_setup(total, dev)      # multi-GPU: parameters and gradients in NVLink peer memory
        self.state = torch.zeros(native.OPT_STATE_WORDS, dtype=torch.int32, device=dev)
        check(lib().mmu_opt_state_init(ptr(self.state), stream()), "mmu_opt_state_init")
        if seed is None:
            # consume one draw from the global generator so torch.manual_seed controls the device stream too
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if self.sample_stream == "device" else 0
        self.seed = D.same_on_all_ranks(int(seed))
        self.mods, off = [], 0
        for i, (e, g) in enumerate(zip(embeds, graphs)):
            self.mods.append(_Modality(e, g if isinstance(g, Graph) else Graph.from_sparse_coo(g), batch_size,
                                       None if refs is None else refs[i], self.flat, off, i, self.seed,
                                       self.sample_stream == "host"))
            off += sizes[i]
            # 1/n_batches of the loss (model.py:453).  A rank that optimises its own block of query rows on its own
            # (sharded transform) still normalises by the batch count of the WHOLE query set.
            self.mods[-1].n_batches_norm = int(norm_batches[i]) if norm_batches is not None else self.mods[-1].n_batches
        # approximate ex2/lg2/rcp force arithmetic only where the stream is not the reference's anyway
        self.fast_math = os.environ.get("MMUMAP_FAST_MATH", "1" if self.sample_stream == "device" else "0") == "1"
        self.peer_tail = os.environ.get("MMUMAP_PEER_TAIL", "auto")      # read once per optimiser, not per epoch
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev) if track_loss else None
        self.losses: list[float] = []
        self.done = 0                  # epochs completed (== the device-side epoch counter)
        self.edge_updates = 0          # host-stream mode counts them exactly; device mode reads kept_count

    def _peer_setup(self, total: int, dev):
        """Multi-GPU (NCCL backend, one node): put the replicated parameter buffer and this rank's partial
        gradient buffer into symmetric memory that every rank maps, so that the epoch's exchange is one kernel
        over peer pointers (mmu_adam_step_peer) instead of an NCCL all-reduce plus a replicated Adam step.
        Returns None (-> NCCL path) for one rank, other backends, MMUMAP_PEER_ADAM=0, or when the symmetric
        allocation is not available on this system."""
        import torch.distributed as dist
        w = D.world()
        if (w == 1 or w > native.PEER_MAX or total == 0 or os.environ.get("MMUMAP_PEER_ADAM", "1") != "1"
                or dist.get_backend() != "nccl"):
            return None
        import ctypes
        pr = _PEER_CACHE.get("buf")
        if pr is False:                                                # tried before, not available
            return None
        if pr is None or pr["cap"] < total:
            # one symmetric allocation per process, reused by later optimisers (fit, then transform): the
            # rendezvous exchanges memory handles through the store and is not free
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            sym = hdl = None
            cap = -(-total // 65536) * 65536
            try:
                import torch.distributed._symmetric_memory as symm_mem
                n_flags = 2 * native.PEER_MAX                          # uint32 flags, kept in the tail of the buffer
                # [parameters | partial gradients | inbox (push exchange: W slots of ceil(cap / W) floats) | flags]
                inbox_floats = cap + 64 * native.PEER_MAX
                sym = symm_mem.empty(2 * cap + inbox_floats + n_flags, dtype=torch.float32, device=dev)
                hdl = symm_mem.rendezvous(sym, dist.group.WORLD)
                sym.zero_()
            except Exception as exc:                                   # noqa: BLE001 - any failure means "not available"
                if D.rank() == 0:
                    print(f"umap_b200: symmetric memory unavailable ({type(exc).__name__}: {exc}); using NCCL all-reduce",
                          flush=True)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)                  # all ranks take the same path
            if int(ok.item()) == 0:
                _PEER_CACHE["buf"] = False
                return None
            torch.cuda.synchronize()
            dist.barrier()                                             # every rank's flags are zero before the first use
            bases = [int(b) for b in hdl.buffer_ptrs]
            arr = ctypes.c_uint64 * w
            # NVSwitch multicast mapping of the same buffer (0 when the system has none): multimem.ld_reduce / multimem.st
            mc = 0
            if os.environ.get("MMUMAP_PEER_MULTIMEM", "1") == "1":
                try:
                    mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                except Exception:                                      # noqa: BLE001
                    mc = 0
            mc_ok = torch.tensor([1 if mc else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(mc_ok, op=dist.ReduceOp.MIN)               # all ranks take the same kernel form
            if int(mc_ok.item()) == 0:
                mc = 0
            pr = {"sym": sym, "hdl": hdl, "seq": 0, "cap": cap,
                  "params": arr(*bases), "grads": arr(*[b + 4 * cap for b in bases]),
                  "inbox": arr(*[b + 8 * cap for b in bases]),
                  "flags": arr(*[b + 4 * (2 * cap + inbox_floats) for b in bases]),
                  "mc_params": mc, "mc_grads": (mc + 4 * cap) if mc else 0,
                  "done": torch.zeros(2, dtype=torch.int32, device=dev)}
            _PEER_CACHE["buf"] = pr
        sym, cap = pr["sym"], pr["cap"]
        # the previous user's last epoch ended with the slot-1 barrier: no peer still reads these buffers
        sym[cap:cap + total].zero_()
        self.flat = (sym[:total], sym[cap:cap + total], self.flat[2], self.flat[3])
        return pr

    # ------------------------------------------------------------------ one epoch
    def _window_rows(self, mod: _Modality) -> int:
        """Rows per tail window of mmu_edge_forces (0 = one pass).  Tables that do not fit the L2 (p and, in fit
        mode, g of the tail side: 10M x 2-D is 80 MB each) are processed in windows of ~48 MB so that the random
        gathers / reds of a pass stay L2 resident; option sgd_window_mb: -1 automatic, 0 never, >0 window bytes."""
        opt = native.get_option("sgd_window_mb")
        if opt == 0 or self.mode == "invert":
            return 0
        row_bytes = mod.dim * 4 * (1 if self.mode == "transform" else 2)
        table = mod.rep_count * row_bytes
        # automatic: only where a row is smaller than a 32-byte DRAM sector (d <= 4: every random access would move
        # 2-4x its payload) AND the tables overflow the L2.  Measured on one B200, 10M x 2-D (80 + 80 MB), ms/epoch:
        # one pass 22.0, windows of 32 MB 18.1, 48 MB 15.3, 64 MB 12.7, 80 MB (two windows of 40 + 40 MB) 10.5 -- the
        # largest slice that still stays L2 resident wins, every extra pass re-reads the records and re-draws the
        # negatives; 1M x 16-D (64-byte rows, no sector waste): 1.5 -> 3.1 ms/epoch, i.e. a loss.
        if opt < 0 and (table <= (100 << 20) or mod.dim * 4 > 16):
            return 0
        window = (AUTO_WINDOW_MB << 20) if opt < 0 else (opt << 20)
        n_win = max(1, -(-table // window))
        if n_win == 1:
            return 0
        rows = -(-mod.rep_count // n_win)
        return -(-rows // 1024) * 1024

    def _forces(self, mod: _Modality, kept_rec, kept_hdr, neg, batch_kept):
        if self.mode == "invert":                                        # model.py:437,447
            mi = self.mods.index(mod)
            check(lib().mmu_invert_forces(ptr(kept_rec), ptr(kept_hdr), ptr(neg), ptr(batch_kept), mod.n_batches_norm,
                                          self.num_rep, mod.rep_count, ptr(mod.p), ptr(mod.ref), ptr(self.sigmas[mi]),
                                          ptr(self.rhos[mi]), ptr(mod.g), mod.dim, self.a, self.b, mod.seed,
                                          ptr(self.state), ptr(self.loss), stream()), "mmu_invert_forces")
            return
        tail = mod.ref if self.mode == "transform" else mod.p
        grad_tail = None if self.mode == "transform" else mod.g
        if profiler.enabled() and not getattr(self, "_capturing", False):
            # per kept edge (2+R) row reads of d*4 B and as many row accumulations in fit mode
            # (one, the query row, in transform mode), + 12 B of edge indices
            rows_touched = (2 + self.num_rep) * 2 if grad_tail is not None else (2 + self.num_rep) + 1
            nbytes = mod.expected_kept * (rows_touched * mod.dim * 4 + 12)
            with profiler.stage("edge_forces", bytes=nbytes, edge_updates=mod.expected_kept * (1 + self.num_rep)):
                self._launch_forces(mod, kept_rec, kept_hdr, neg, batch_kept, tail, grad_tail)
            return
        self._launch_forces(mod, kept_rec, kept_hdr, neg, batch_kept, tail, grad_tail)

    def _launch_forces(self, mod, kept_rec, kept_hdr, neg, batch_kept, tail, grad_tail):
        if mod.window_rows is None:
            mod.window_rows = self._window_rows(mod)
        check(lib().mmu_edge_forces(ptr(kept_rec), ptr(kept_hdr), ptr(neg), ptr(batch_kept), mod.n_batches_norm, self.num_rep,
                                    mod.rep_count, ptr(mod.p), ptr(tail), ptr(mod.g), ptr(grad_tail), mod.dim, self.a,
                                    self.b, mod.seed, ptr(self.state), ptr(self.loss), int(self.fast_math),
                                    mod.window_rows, stream()), "mmu_edge_forces")

    def _infonce(self, src: _Modality, dst: _Modality, perm, neg, stream_id: int):
        num = min(src.count, dst.count)
        a_lo, a_hi = D.item_range(num, D.rank(), D.world())
        check(lib().mmu_infonce_range(ptr(src.p), ptr(dst.p), num, a_lo, a_hi, src.dim, ptr(perm), ptr(neg), INFONCE_NEG,
                                      INFONCE_CHUNK, self.alpha, INFONCE_TAU, ptr(src.g), ptr(dst.g), self.seed,
                                      stream_id, ptr(self.state), ptr(self.loss), stream()), "mmu_infonce_range")

    def epoch(self):
        host = self.sample_stream == "host"
        dev = torch.device("cuda")
        for mod in self.mods:
            if host:
                kept, neg, counts = replay_host_draws(mod, self.num_rep)
                self.edge_updates += int(kept.numel()) * (1 + self.num_rep)
                n = mod.load_host_draws(kept, counts)
                neg_d = neg.pin_memory().to(dev, non_blocking=True) if n else torch.zeros(1, dtype=torch.int32, device=dev)
                self._forces(mod, mod.kept_rec, mod.kept_hdr, neg_d, mod.batch_kept)
            else:
                g = mod.graph
                with profiler.stage("edge_sample", level=2):
                    check(lib().mmu_edge_sample_range(ptr(g.row), ptr(g.col), ptr(g.val), mod.e_lo, mod.e_hi,
                                                      mod.batch_size, mod.n_batches, mod.seed, ptr(self.state),
                                                      ptr(mod.kept_rec), ptr(mod.kept_hdr), ptr(mod.batch_kept),
                                                      stream()), "mmu_edge_sample_range")
                self._forces(mod, mod.kept_rec, mod.kept_hdr, None, mod.batch_kept)
        self._epoch_tail()

    def _epoch_tail(self):
        """InfoNCE, gradient all-reduce, Adam: everything of an epoch after the force kernels."""
        self._infonce_all(stream())
        self._adam_tail()

    def _infonce_all(self, st):
        """Cross-modal InfoNCE gradients of one epoch, launched on stream `st` (model.py:459-472)."""
        host = self.sample_stream == "host"
        dev = torch.device("cuda")
        if self.mode == "fit":
            n = len(self.mods)
            sid = 0
            for i in range(n):
                for j in range(i + 1, n):
                    src, dst = self.mods[i], self.mods[j]
                    num = min(src.count, dst.count)
                    if num == 0:
                        continue
                    a_lo, a_hi = D.item_range(num, D.rank(), D.world())
                    pf = nf = pr = nr = None
                    if host:
                        # the reference draws (randperm, randint...) for L_ij and then for L_ji (model.py:463-466)
                        (pf, nf), (pr, nr) = replay_infonce_draws(num), replay_infonce_draws(num)
                        pf, nf, pr, nr = (t.pin_memory().to(dev, non_blocking=True) for t in (pf, nf, pr, nr))
                    # both directions in one grid: they read the same embedding state
                    with profiler.stage("infonce", level=2):
                        check(lib().mmu_infonce_bidir(ptr(src.p), ptr(dst.p), num, a_lo, a_hi, src.dim, ptr(pf), ptr(nf),
                                                      ptr(pr), ptr(nr), INFONCE_NEG, INFONCE_CHUNK, self.alpha, INFONCE_TAU,
                                                      ptr(src.g), ptr(dst.g), self.seed, sid, ptr(self.state),
                                                      ptr(self.loss), st), "mmu_infonce_bidir")
                    sid += 2

    def _adam_tail(self):
        p, g, m, v = self.flat
        if self.peer is not None:
            # multi-GPU over peer memory: [all gradients complete] -> reduce my shard + Adam + write all replicas
            # -> [all parameters delivered] -> clear my gradient buffer
            pr, w, r = self.peer, D.world(), D.rank()
            pr["seq"] += 1
            tail = self.peer_tail
            if tail == "auto":
                # small tables: NVLink latency is the cost -> push form (two one-way hops); large tables: the pull form,
                # whose in-switch reduction (multimem) moves half the bytes
                tail = "push" if self.total * 4 <= PUSH_TAIL_MAX_BYTES else "fused"
            if tail == "push":
                slot = (-(-(self.total // 4) // w) + 1) * 4
                with profiler.stage("epoch_tail", level=2):
                    check(lib().mmu_epoch_tail_push(pr["params"], pr["inbox"], pr["flags"], ptr(g), ptr(m), ptr(v), self.total,
                                                    slot, w, r, self.lr, BETA1, BETA2, EPS, ptr(self.state), ptr(pr["done"]),
                                                    stream()), "mmu_epoch_tail_push")
                self.done += 1
                if self.loss is not None:
                    D.all_reduce_sum(self.loss)
                    self.losses.append(float(self.loss.item()))
                    self.loss.zero_()
                return
            if tail == "fused":
                # ONE launch: barrier + shard reduce + Adam + replica store + gradient clear + barrier + state advance
                # (barrier sequence number on the device, pr["done"][1]: no per-epoch argument -> graph replayable;
                #  pr["seq"] mirrors it on the host for the legacy barrier calls)
                with profiler.stage("epoch_tail", level=2):
                    check(lib().mmu_epoch_tail_peer(pr["params"], pr["grads"], pr["flags"], pr["mc_params"], pr["mc_grads"],
                                                    ptr(m), ptr(v), self.total, w, r, 0, self.lr, BETA1, BETA2, EPS,
                                                    ptr(self.state), ptr(pr["done"]), stream()), "mmu_epoch_tail_peer")
                self.done += 1
                if self.loss is not None:
                    D.all_reduce_sum(self.loss)
                    self.losses.append(float(self.loss.item()))
                    self.loss.zero_()
                return
            check(lib().mmu_opt_state_advance(ptr(self.state), self.lr, BETA1, BETA2, stream()), "mmu_opt_state_advance")
            with profiler.stage("adam", level=2):
                check(lib().mmu_peer_barrier(pr["flags"], w, r, 0, pr["seq"], stream()), "mmu_peer_barrier")
                check(lib().mmu_adam_step_peer(pr["params"], pr["grads"], ptr(m), ptr(v), self.total, w, r, BETA1, BETA2,
                                               EPS, ptr(self.state), stream()), "mmu_adam_step_peer")
                check(lib().mmu_peer_barrier(pr["flags"], w, r, 1, pr["seq"], stream()), "mmu_peer_barrier")
                g.zero_()
        else:
            # (NCCL path) one all-reduce of the flat gradient buffer, then the identical Adam step everywhere
            D.all_reduce_sum(g)
            check(lib().mmu_opt_state_advance(ptr(self.state), self.lr, BETA1, BETA2, stream()), "mmu_opt_state_advance")
            with profiler.stage("adam", level=2):
                check(lib().mmu_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), self.total, BETA1, BETA2, EPS, ptr(self.state), 1,
                                          stream()), "mmu_adam_step")
        self.done += 1
        if self.loss is not None:
            D.all_reduce_sum(self.loss)
            self.losses.append(float(self.loss.item()))
            self.loss.zero_()

    def _run_overlapped(self, epochs: int):
        """Device sample stream, large problems: the Bernoulli sampling of epoch e+1 does not depend on the
        embeddings, so it runs on a second stream (double-buffered kept lists, explicit epoch number) while
        the force kernels of epoch e execute; the main stream only waits for the sampled-event."""
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream()                 # InfoNCE beside the force kernels
        samp = torch.cuda.Stream()                 # the next epoch's sampling (its own branch: not queued behind InfoNCE)
        side.wait_stream(main)
        samp.wait_stream(main)
        dev = torch.device("cuda")
        bufs = []
        for mod in self.mods:
            bufs.append([(mod.kept_rec, mod.kept_hdr, mod.batch_kept),
                         (torch.empty_like(mod.kept_rec), mod._new_hdr(), torch.zeros_like(mod.batch_kept))])
            mod.all_hdrs = [b[1] for b in bufs[-1]]
        sampled = [None, None]
        consumed = [None, None]
        base = self.done
        # The InfoNCE kernel (latency bound, 2 blocks per SM) reads the same embedding state as the force kernels
        # (L1->crossbar bound) and only adds to the gradient buffer, so it runs beside them on the second stream:
        # after the previous epoch's Adam step, before this epoch's.
        side_nce = self.mode == "fit" and len(self.mods) > 1 and os.environ.get("MMUMAP_NCE_OVERLAP", "1") == "1"
        stepped = None

        def issue_sample(e):
            b = e & 1
            with torch.cuda.stream(samp):
                if consumed[b] is not None:
                    samp.wait_event(consumed[b])
                for mi, mod in enumerate(self.mods):
                    g = mod.graph
                    kp, kc, bk = bufs[mi][b]
                    check(lib().mmu_edge_sample_at(ptr(g.row), ptr(g.col), ptr(g.val), mod.e_lo, mod.e_hi, mod.batch_size,
                                                   mod.n_batches, mod.seed, base + e, ptr(self.state), ptr(kp), ptr(kc),
                                                   ptr(bk), samp.cuda_stream), "mmu_edge_sample_at")
                ev = torch.cuda.Event()
                ev.record(samp)
                sampled[b] = ev

        issue_sample(0)
        for e in range(epochs):
            b = e & 1
            nce_done = None
            if side_nce:
                with torch.cuda.stream(side):
                    if stepped is not None:
                        side.wait_event(stepped)
                    self._infonce_all(side.cuda_stream)
                    nce_done = torch.cuda.Event()
                    nce_done.record(side)
            if e + 1 < epochs:
                issue_sample(e + 1)
            main.wait_event(sampled[b])
            for mi, mod in enumerate(self.mods):
                kp, kc, bk = bufs[mi][b]
                mod.kept_hdr = kc                        # kept_last_epoch() reads the buffer in use
                self._forces(mod, kp, kc, None, bk)
            ev = torch.cuda.Event()
            ev.record(main)
            consumed[b] = ev
            if side_nce:
                main.wait_event(nce_done)
            else:
                self._infonce_all(stream())
            self._adam_tail()
            if side_nce:
                stepped = torch.cuda.Event()
                stepped.record(main)
        main.wait_stream(side)
        main.wait_stream(samp)

    def _run_graphed(self, epochs: int):
        """Multi-GPU (fused peer epoch tail), device sample stream: at 8 GPUs an epoch of BASELINE.json configs[1] is ~65 us
        of GPU time but ~8 launches + 6 event operations of host time -- the host, not the GPU, sets the pace.  The whole
        epoch {InfoNCE and the NEXT epoch's sampling on a forked branch | this epoch's force kernels} -> fused tail is
        therefore captured ONCE per kept-list parity (the lists are double buffered) into two CUDA graphs and replayed:
        one graph launch per epoch.  Every per-epoch scalar lives on the device (epoch counter and Adam step in the
        optimiser state, the barrier sequence beside the completion counter), so the graphs have no arguments."""
        eager = 2
        self._run_overlapped(eager)                               # loads every kernel before capture, advances the state
        epochs -= eager
        main = torch.cuda.current_stream()
        cap = torch.cuda.Stream()
        side = torch.cuda.Stream()
        samp = torch.cuda.Stream()
        bufs = []
        for mod in self.mods:
            bufs.append([(mod.kept_rec, mod.kept_hdr, mod.batch_kept),
                         (torch.empty_like(mod.kept_rec), mod._new_hdr(), torch.zeros_like(mod.batch_kept))])
            mod.all_hdrs = list(getattr(mod, "all_hdrs", [])) + [b[1] for b in bufs[-1]]

        def sample_into(b, epoch, st):
            for mi, mod in enumerate(self.mods):
                g = mod.graph
                kp, kc, bk = bufs[mi][b]
                check(lib().mmu_edge_sample_at(ptr(g.row), ptr(g.col), ptr(g.val), mod.e_lo, mod.e_hi, mod.batch_size,
                                               mod.n_batches, mod.seed, epoch, ptr(self.state), ptr(kp), ptr(kc), ptr(bk), st),
                      "mmu_edge_sample_at")

        sample_into(0, self.done, main.cuda_stream)               # the first graphed epoch's list, explicit epoch number
        fit_nce = self.mode == "fit" and len(self.mods) > 1
        graphs, per_epoch = [], 0
        cap.wait_stream(main)
        side.wait_stream(main)
        samp.wait_stream(main)
        self._capturing = True
        try:
            for b in (0, 1):
                gph = torch.cuda.CUDAGraph()
                before = lib().mmu_launch_count()
                with torch.cuda.stream(cap):
                    gph.capture_begin()
                    try:
                        fork = torch.cuda.Event()
                        fork.record(cap)
                        side.wait_event(fork)
                        samp.wait_event(fork)
                        with torch.cuda.stream(side):
                            if fit_nce:
                                self._infonce_all(side.cuda_stream)
                            join = torch.cuda.Event()
                            join.record(side)
                        with torch.cuda.stream(samp):
                            sample_into(b ^ 1, -2, samp.cuda_stream)          # next epoch = device counter + 1
                            join2 = torch.cuda.Event()
                            join2.record(samp)
                        for mi, mod in enumerate(self.mods):
                            kp, kc, bk = bufs[mi][b]
                            self._forces(mod, kp, kc, None, bk)
                        cap.wait_event(join)
                        cap.wait_event(join2)
                        self._adam_tail()
                    finally:
                        gph.capture_end()
                per_epoch = lib().mmu_launch_count() - before
                self.done -= 1                                    # the capture pass counted itself without executing
                self.peer["seq"] -= 1
                graphs.append(gph)
        finally:
            self._capturing = False
        main.wait_stream(cap)
        for e in range(epochs):
            graphs[e & 1].replay()
        lib().mmu_launch_count_add(max(epochs - 2, 0) * per_epoch)    # the two capture passes counted themselves once
        self.done += epochs
        self.peer["seq"] += epochs
        for mi, mod in enumerate(self.mods):
            mod.kept_hdr = bufs[mi][(epochs - 1) & 1][1] if epochs > 0 else mod.kept_hdr
        self._graphs = graphs                                     # keep alive until the stream has drained

    def run(self, epochs: int):
        """`epochs` optimiser epochs.  With the device sample stream an epoch is a fixed sequence of
        launches whose only varying inputs (epoch counter, Adam step) live in device memory, so it
        is captured once into a CUDA graph and replayed: one graph launch per epoch instead of a
        dozen kernel launches (sub-millisecond epochs are launch-latency territory)."""
        mode = os.environ.get("MMUMAP_GRAPH", "auto")
        small = sum(m.graph.nnz for m in self.mods) <= 1_000_000        # epochs of a few tens of microseconds
        use_graph = (self.sample_stream == "device" and self.loss is None and epochs > 2
                     and (mode == "1" or (mode == "auto" and small))
                     and (D.world() == 1 or (self.peer is None and os.environ.get("MMUMAP_GRAPH_NCCL", "0") == "1")))
        peer_graph = (self.peer is not None and self.sample_stream == "device" and self.loss is None and epochs >= 12
                      and self.mode in ("fit", "transform") and not profiler.enabled(2)
                      and self.peer_tail in ("auto", "fused", "push")
                      and os.environ.get("MMUMAP_EPOCH_GRAPH", "1") == "1")
        if peer_graph:
            self._run_graphed(epochs)
            return self.result()
        if not use_graph:
            overlap = (self.sample_stream == "device" and self.mode in ("fit", "transform")
                       and epochs > 1 and os.environ.get("MMUMAP_OVERLAP_SAMPLE", "1") == "1")
            if overlap:
                self._run_overlapped(epochs)
            else:
                for _ in range(epochs):
                    self.epoch()
            return self.result()
        start = self.done
        self.epoch()                                   # eager first epoch: loads every kernel before capture
        graph = torch.cuda.CUDAGraph()
        before = lib().mmu_launch_count()
        # plain capture_begin/end on a side stream: the torch.cuda.graph() context manager also runs
        # gc.collect() and empty_cache(), which costs more than the launches it saves
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            graph.capture_begin()
            try:
                self.epoch()                           # recorded, not executed
            finally:
                graph.capture_end()
        cur.wait_stream(side)
        per_epoch = lib().mmu_launch_count() - before
        lib().mmu_launch_count_add((epochs - 2) * per_epoch)      # the capture pass itself counted once
        for _ in range(epochs - 1):
            graph.replay()
        self.done = start + epochs                     # the capture pass counted itself without executing
        self._graph = graph                            # keep alive until the stream has drained
        return self.result()

    def result(self):
        out = [m.p.clone() for m in self.mods]
        # one read of the overflow flags per run (the clone above already orders after the last epoch)
        flags = torch.stack([h[2] for m in self.mods for h in getattr(m, "all_hdrs", [m.kept_hdr])])
        if bool(flags.any().item()):
            raise native.NativeError("kept-edge list overflow: an epoch kept more edges than sum(w) + 10 sigma")
        return out

    def kept_last_epoch(self) -> int:
        """Kept edges of the last epoch over all ranks."""
        t = torch.stack([m.kept_hdr[0] for m in self.mods]).sum().to(torch.int64).reshape(1)
        D.all_reduce_sum(t)
        return int(t.item())
