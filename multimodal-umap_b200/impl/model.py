"""Drop-in replacement for the reference's impl/model.py on a B200.

Same public surface as /root/reference/impl/model.py (device, UMAPEncoder, UMAPMixture with
fit / fit_transform / transform / inverse_transform / save_state_dict / load_state_dict) so that
the reference's main.py, impl/validation.py and impl/crossmodal.py run unchanged when this
directory precedes the reference tree on sys.path (`impl` is a namespace package: there is
deliberately no impl/__init__.py).  All numerics run in hand-written sm_100a kernels behind
the C ABI of include/mmumap.h; there is no CPU fallback.

Engine switches (attributes of UMAPMixture / environment):
  sample_stream  "device" (Philox in-kernel, default) | "host" (replay the reference's CPU
                 generator draws; parity mode)                        env MMUMAP_SAMPLE_STREAM
  sigma_solver   "bisect" (default) | "newton" (the reference's 20-step Newton, reproducing
                 its divergent rows)                                  env MMUMAP_SIGMA
  knn_method     "tc" (tcgen05 candidates + fp32 rescoring) | "simt"  env MMUMAP_KNN
"""
from __future__ import annotations

import functools
import math
import os
import weakref

import torch

from umap_b200 import graph as G
from umap_b200 import dist as D
from umap_b200 import native, profiler
from umap_b200.layout import LayoutOptimizer
from umap_b200.spectral import spectral_init

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")      # ref: model.py:10

_GRAPHS: dict[int, tuple] = {}     # id(sparse tensor) -> (weakref, Graph); keeps checkpoints torch-only


def _register(t: torch.Tensor, g: G.Graph) -> torch.Tensor:
    _GRAPHS[id(t)] = (weakref.ref(t, lambda _r, k=id(t): _GRAPHS.pop(k, None)), g)
    return t


def _as_graph(t) -> G.Graph:
    if isinstance(t, G.Graph):
        return t
    hit = _GRAPHS.get(id(t))
    if hit is not None and hit[0]() is t:
        return hit[1]
    g = G.Graph.from_sparse_coo(t)
    _register(t, g)
    return g


def _coo(g: G.Graph) -> torch.Tensor:
    idx = torch.stack([g.row.long(), g.col.long()], dim=0)
    t = torch.sparse_coo_tensor(idx, g.val, (g.n_rows, g.n_cols), is_coalesced=True)
    return _register(t, g)


DEFAULT_KNN_DIST = "rows"      # multi-GPU kNN: query row-blocks per rank against the device-resident database


class _Uploads:
    """Device copies of a list of inputs.  Host tensors are copied on a side stream in list order, so the
    upload of modality i+1 overlaps the graph construction of modality i (pinned host memory makes the
    copies asynchronous); indexing waits for that tensor's copy only.

    Multi-GPU (one process per GPU): a rank uploads only ITS row block of a host tensor (1/W of the bytes over its
    host link, the reference's single x.to(device) of model.py:634 split W ways) and the full matrix is assembled
    on every GPU by one all-gather, device to device over NVLink -- instead of W ranks each pulling the whole input
    through the host links."""

    _side = {}                                     # device index -> the one upload stream of this process

    def __init__(self, inputs, shard: bool = False):
        native.require_cuda()                      # no CPU fallback: fail before touching any stream
        self._items = []
        self.h2d_bytes = 0
        w, r = D.world(), D.rank()
        cur = torch.cuda.current_stream()
        # Destinations come from the CURRENT stream's pool, all of them before any copy is issued: a block allocated
        # under a side stream belongs to that stream's pool and is never handed to the next fit (measured: +0.94 GiB
        # reserved and two cudaMalloc calls per C2 fit, without bound).  The side stream then waits for the current
        # one once (whatever last used these blocks was queued there) and only carries the copies.
        plan = []
        for x in inputs:
            if x.is_cuda:
                plan.append((x, None, None, None))
                continue
            if shard and w > 1 and x.dim() == 2 and x.shape[0] >= w:
                n = x.shape[0]
                lo, hi = D.row_block(n, r, w)
                per = D.block_size(n, w)
                blk = torch.empty((per, x.shape[1]), dtype=x.dtype, device=device)
                if hi - lo < per:
                    blk[hi - lo:].zero_()
                plan.append((blk, x[lo:hi], hi - lo, n))
            else:
                plan.append((torch.empty(x.shape, dtype=x.dtype, device=device), x, None, None))
        side = None
        for t, src, rows, gather in plan:
            if src is None:
                self._items.append((t, None, None))
                continue
            if side is None:
                key = torch.cuda.current_device()
                side = _Uploads._side.get(key)
                if side is None:
                    side = _Uploads._side[key] = torch.cuda.Stream()
                side.wait_stream(cur)
            with torch.cuda.stream(side):
                if rows is None:
                    t.copy_(src, non_blocking=True)
                elif rows > 0:
                    t[:rows].copy_(src, non_blocking=True)
                self.h2d_bytes += src.numel() * src.element_size()
                ev = torch.cuda.Event()
                ev.record(side)
            self._items.append((t, ev, gather))

    def __len__(self):
        return len(self._items)

    def __del__(self):
        # a copy nobody waited for (a fit that raised half way): order it before whatever the current stream does
        # next, so that its destination block is not reused under it
        try:
            for _, ev, _ in self._items:
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
        except Exception:
            pass

    def __getitem__(self, i):
        t, ev, gather = self._items[i]
        if ev is not None:
            # after this wait every use of `t` is ordered on the current stream, the one its block was allocated
            # under: freeing it needs no record_stream bookkeeping
            torch.cuda.current_stream().wait_event(ev)
            if gather is not None:
                import torch.distributed as dist
                full = torch.empty((D.world() * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)
                dist.all_gather_into_tensor(full, t)          # every rank indexes the inputs in the same order
                t = full[:gather]
            self._items[i] = (t, None, None)
        return t


@functools.lru_cache(maxsize=64)
def _ab_coeffs(min_dist: float, num_iters: int):
    """Gauss-Newton fit of model.py:587-618 (a pure function of min_dist: cached per value)."""
    x = torch.linspace(1e-4, 3.0, 200, dtype=torch.float32)
    target = torch.where(x <= min_dist, torch.tensor(1.0), torch.exp(-(x - min_dist)))
    betas = torch.tensor([1.0, 1.0])
    for _ in range(num_iters):
        a_, b_ = betas[0].abs() + 1e-6, betas[1].abs() + 1e-6
        xp = x.pow(2 * b_)
        est = 1.0 / (1.0 + a_ * xp)
        res = target - est
        jac = torch.stack([torch.sign(betas[0]) * xp * est * est,
                           torch.sign(betas[1]) * a_ * xp * 2.0 * torch.log(x) * est * est], dim=1)
        betas = betas - torch.linalg.pinv(jac) @ res
    return (betas[0].abs() + 1e-6).item(), (betas[1].abs() + 1e-6).item()


class UMAPEncoder:
    """Single-modality graph builder and initialiser (ref: model.py:12-278)."""

    def __init__(self, k_neighbors: int, out_dim: int, id: int = 0):
        self.k_neighbors = k_neighbors
        self.out_dim = out_dim
        self.id = id
        self.sigmas = None
        self.rhos = None
        self.sigma_solver = os.environ.get("MMUMAP_SIGMA", "bisect")
        self.knn_method = None

    def get_sigmas(self, dists: torch.Tensor, min_dists: torch.Tensor, num_iters: int = 20) -> torch.Tensor:
        """ref: model.py:33-61.  sigma_i with sum_j exp(-(d_ij - rho_i)/sigma_i) = log2(k).  `min_dists` is rho
        repeated along the row, as the reference passes it (model.py:199-200); the kernel takes rho as the row
        minimum itself, so anything else is rejected rather than silently ignored.  `num_iters` is the Newton
        iteration count (solver "newton"); the bisection solver always runs its 64 halvings."""
        dists = dists.to("cuda", torch.float32).reshape(-1, self.k_neighbors).contiguous()
        if min_dists is not None:
            md = torch.as_tensor(min_dists, dtype=torch.float32, device=dists.device).reshape(dists.shape[0], -1)
            if not torch.equal(md.amin(dim=1), dists.amin(dim=1)) or not torch.equal(md.amin(dim=1), md.amax(dim=1)):
                raise ValueError("get_sigmas: min_dists must be the row minimum of dists (rho), as in model.py:199")
        idx = torch.arange(self.k_neighbors, dtype=torch.int32, device=dists.device).repeat(dists.shape[0], 1)
        n_iter = int(num_iters) if self.sigma_solver == "newton" else max(64, int(num_iters))
        _, _, sigma, _ = G.smooth_knn(idx, dists, self.sigma_solver, n_iter)
        return sigma

    def fuzzy_knn_graph(self, inputs: torch.Tensor, mode: str = "fit", query: torch.Tensor | None = None,
                        ref_data: torch.Tensor | None = None, num_iters: int = 10, a: float | None = None,
                        b: float | None = None) -> torch.Tensor:
        """ref: model.py:63-209.  Exact kNN (self excluded iff ref_data is None, model.py:87-90),
        then rho/sigma/weights (or 1/(1+a d^2b) in invert mode); returns the coalesced (Q, N) COO."""
        native.require_cuda()
        q = inputs if query is None else query
        idx, dist = G.knn_graph(q, inputs, self.k_neighbors, exclude_self=ref_data is None, method=self.knn_method)
        if mode != "invert":
            with profiler.stage("smooth_knn", bytes=float(idx.shape[0]) * (16 * self.k_neighbors + 8)):
                col, w, sigma, rho = G.smooth_knn(idx, dist, self.sigma_solver)
            if mode == "fit":
                self.sigmas, self.rhos = sigma, rho                      # model.py:202-204
        else:
            col, w = G.invert_weights(idx, dist, a, b)
        g = G.Graph.from_fixed_degree(col, w, inputs.shape[0])
        g.k = self.k_neighbors
        g.col2d, g.w2d = col, w
        return _coo(g)

    @torch.no_grad()
    def embed_all(self, input: torch.Tensor) -> torch.Tensor:
        """ref: model.py:211-234."""
        return spectral_init(_as_graph(input), self.out_dim)

    @torch.no_grad()
    def embed_query(self, ref: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
        """ref: model.py:236-252."""
        g = _as_graph(query)
        if getattr(g, "col2d", None) is not None:
            return G.embed_query(g.col2d, g.w2d, ref)
        rs = torch.zeros(g.n_rows, dtype=torch.float32, device=g.val.device)
        rs.index_add_(0, g.row.long(), g.val)
        return G.spmm(g, ref.to("cuda", torch.float32), g.val / rs.clamp(min=1e-6)[g.row.long()])

    def init(self, input: torch.Tensor, mode: str = "fit", query: torch.Tensor | None = None,
             ref_data: torch.Tensor | None = None, ref_embeds: torch.Tensor | None = None,
             a: float | None = None, b: float | None = None):
        """ref: model.py:254-278."""
        graph = self.fuzzy_knn_graph(input, mode, query, ref_data, num_iters=10, a=a, b=b)
        if mode == "fit":
            g = _as_graph(graph)
            with profiler.stage("fuzzy_union"):
                sym = G.fuzzy_union(g.col2d, g.w2d)                      # model.py:271
            graph = _coo(sym)
            with profiler.stage("spectral_init"):
                # multi-GPU: modality m is solved by rank m % world and broadcast (embeddings are replicated)
                owner = self.id % D.world()
                if D.rank() == owner:
                    embed = self.embed_all(graph).contiguous()
                else:
                    embed = torch.empty((input.shape[0], self.out_dim), dtype=torch.float32, device="cuda")
                D.broadcast(embed, owner)
        elif mode == "transform":
            embed = self.embed_query(ref_embeds, graph)
        else:
            embed = self.embed_query(ref_embeds if ref_embeds is not None else input, graph)
        return graph, embed


class UMAPMixture:
    """Multimodal UMAP with InfoNCE alignment (ref: model.py:280-714)."""

    def __init__(self, k_neighbors: int, out_dim: int, min_dist: float, num_encoders: int):
        self.k_neighbors = k_neighbors
        self.out_dim = out_dim
        self.min_dist = min_dist
        self.num_encoders = num_encoders
        self.a, self.b = self.get_ab_coeffs(min_dist)
        self.encoders = [UMAPEncoder(k_neighbors, out_dim, id=i) for i in range(num_encoders)]
        self.data = None
        self.graphs = []
        self.embeds = []
        self._engine_defaults()

    def _engine_defaults(self):
        self.sample_stream = os.environ.get("MMUMAP_SAMPLE_STREAM", "device")
        self.last_optimizer = None
        self.last_h2d_bytes = 0

    # ------------------------------------------------------------------ optimiser
    def _train(self, embeds, graphs, epochs: int, num_rep: int, lr: float, alpha: float, batch_size: int,
               mode: str = "fit", data_indices: list | None = None, desc: str = "Training"):
        """ref: model.py:396-481."""
        native.require_cuda()
        refs = sigmas = rhos = None
        if mode == "invert":
            # ref: model.py:418-420 (tails are the target modality's data rows) and :437,:447.  The reference
            # indexes self.encoders with the LOOP index there; the target encoder data_indices[i] is what the
            # sigma[j_idx] / rho[j_idx] lookups need and what is used here.
            n_modes = len(embeds) if data_indices is None else len(data_indices)
            tgt = [data_indices[i] if data_indices is not None else i for i in range(n_modes)]
            refs = [self.data[t] for t in tgt]
            sigmas = [self.encoders[t].sigmas for t in tgt]
            rhos = [self.encoders[t].rhos for t in tgt]
        if mode == "transform":
            for ref in self.embeds:                                       # model.py:399-401
                ref.requires_grad = False
            n_modes = len(embeds) if data_indices is None else len(data_indices)
            refs = [self.embeds[data_indices[i]] if data_indices is not None else self.embeds[i]
                    for i in range(n_modes)]
        opt = LayoutOptimizer(embeds, [_as_graph(g) for g in graphs], self.a, self.b, num_rep, lr, alpha,
                              batch_size, mode=mode, refs=refs, sigmas=sigmas, rhos=rhos,
                              sample_stream=getattr(self, "sample_stream", None),
                              seed=getattr(self, "_shard_seed", None), norm_batches=getattr(self, "_shard_norm_batches", None))
        with profiler.stage("optimise", epochs=epochs):
            out = opt.run(epochs)
        self.last_optimizer = opt
        return [e.requires_grad_(True) for e in out]                     # leaves, as model.py:397,481

    def fit(self, inputs: list, epochs: int, num_rep: int = 8, lr: float = 0.2, alpha: float = 0.5,
            batch_size: int = 512) -> None:
        """ref: model.py:483-508.  (The reference uploads the inputs twice, :496 and :634; here the device
        copies made for the graph stage are the ones kept as self.data.)"""
        inputs = _Uploads(inputs, shard=True)
        graphs, embeds = self.init(inputs, mode="fit")
        self.graphs = graphs
        self.data = [inputs[i] for i in range(len(inputs))]
        self.last_h2d_bytes = inputs.h2d_bytes              # what THIS rank pulled over its host link
        self.embeds = self._train(embeds, graphs, epochs, num_rep, lr, alpha, batch_size, mode="fit",
                                  desc=f"Training {self.num_encoders} encoders")

    def fit_transform(self, inputs: list, epochs: int, num_rep: int = 8, lr: float = 0.2, alpha: float = 0.5,
                      batch_size: int = 512):
        """ref: model.py:510-525."""
        self.fit(inputs, epochs, num_rep, lr, alpha, batch_size)
        return self.embeds

    def transform(self, inputs: list, epochs: int, data_indices: list | None = None, num_rep: int = 8,
                  lr: float = 0.2, alpha: float = 0.5, batch_size: int = 512):
        """ref: model.py:527-555."""
        sharded = self._sharded_rows(inputs, batch_size)
        if sharded is not None:
            return self._run_sharded(sharded, "transform", inputs, epochs, data_indices, num_rep, lr, alpha, batch_size)
        graphs, embeds = self.init(inputs, mode="transform", data_indices=data_indices)
        return self._train(embeds, graphs, epochs, num_rep, lr, alpha, batch_size, mode="transform",
                           data_indices=data_indices, desc=f"Embedding {len(embeds)} modalities")

    # ------------------------------------------------------------------ multi-GPU: independent query rows
    def _sharded_rows(self, inputs, batch_size: int):
        """Multi-GPU transform / inverse_transform (device sample stream): the queries are independent -- only the
        query's own row receives gradient (model.py:399-401,416) -- so rank r takes a block of whole row-batches of
        every input and runs the single-GPU path on it with NO exchange; the blocks are all-gathered at the end.
        Returns the per-input (lo, hi, per) blocks, or None when the call should not be sharded."""
        w = D.world()
        if w == 1 or getattr(self, "sample_stream", "device") != "device" or os.environ.get("MMUMAP_SHARD_TRANSFORM", "1") != "1":
            return None
        align = batch_size * 128 // math.gcd(batch_size, 128)          # whole row-batches AND whole 128-row kNN tiles
        blocks = []
        for x in inputs:
            n = x.shape[0] if x.dim() > 1 else 1
            if n < 2 * w * align:
                return None
            lo, hi = D.row_block(n, D.rank(), w, align)
            blocks.append((lo, hi, D.block_size(n, w, align), n))
        return blocks

    def _run_sharded(self, blocks, mode, inputs, epochs, data_indices, num_rep, lr, alpha, batch_size):
        r = D.rank()
        local = [x[lo:hi] for x, (lo, hi, _, _) in zip(inputs, blocks)]
        outs = [None] * len(inputs)
        base = int(torch.randint(0, 2 ** 62, (1,)).item())               # same draw on every rank (same generator state)
        if all(hi > lo for (lo, hi, _, _) in blocks):
            self._shard_seed = (base + r * 0x632BE59BD9B4E019) & 0x3FFFFFFFFFFFFFFF
            self._shard_norm_batches = [-(-n // batch_size) for (_, _, _, n) in blocks]
            try:
                with D.local_only():
                    graphs, embeds = self.init(local, mode=mode, data_indices=data_indices)
                    outs = self._train(embeds, graphs, epochs, num_rep, lr, alpha, batch_size, mode=mode,
                                       data_indices=data_indices, desc=f"{mode} of {len(embeds)} modalities (row block {r})")
            finally:
                self._shard_seed = self._shard_norm_batches = None
        full = []
        for o, (lo, hi, per, n), x in zip(outs, blocks, inputs):
            width = o.shape[1] if o is not None else (self.out_dim if mode == "transform" else self.data[0].shape[1])
            loc = o.detach() if o is not None else torch.zeros((0, width), dtype=torch.float32, device="cuda")
            full.append(D.all_gather_rows(loc, n, per).requires_grad_(True))
        return full

    def inverse_transform(self, inputs: list, epochs: int, data_indices: list | None = None, num_rep: int = 8,
                          lr: float = 0.2, alpha: float = 0.5, batch_size: int = 512):
        """ref: model.py:557-585.  The reference's invert path raises a shape error as shipped
        (its initial value is Q x out_dim instead of Q x D).  The evident
        intent is implemented: kNN of the query embeddings among the fitted embeddings with weights
        1/(1+a d^2b) (model.py:206), initial value = weighted mean of the TARGET modality's data rows
        (Q x D), then `epochs` epochs of the invert-mode losses (model.py:336-362) under Adam."""
        graphs, embeds = self.init(inputs, mode="invert", data_indices=data_indices)
        return self._train(embeds, graphs, epochs, num_rep, lr, alpha, batch_size, mode="invert",
                           data_indices=data_indices, desc=f"Inverting {len(embeds)} modalities")

    # ------------------------------------------------------------------ curve fit (host, one-off)
    def get_ab_coeffs(self, min_dist: float, num_iters: int = 50):
        """ref: model.py:587-618: Gauss-Newton fit of 1/(1+a x^(2b)) to the min_dist target on
        linspace(1e-4, 3, 200) from (1, 1), fp32; closed-form Jacobian instead of autograd."""
        return _ab_coeffs(float(min_dist), int(num_iters))

    # ------------------------------------------------------------------ orchestration
    def init(self, inputs: list, mode: str = "fit", data_indices: list | None = None):
        """ref: model.py:620-651."""
        if mode not in ["fit", "transform", "invert"]:
            raise ValueError(f"Invalid mode: {mode}")
        native.require_cuda()
        if not isinstance(inputs, _Uploads):
            inputs = _Uploads(inputs)
        graphs, embeds = [], []
        encoder_indices = data_indices if data_indices is not None else range(self.num_encoders)
        if mode == "fit" and D.world() > 1:
            return self._init_fit_distributed(inputs, list(encoder_indices))
        for idx, i in enumerate(encoder_indices):
            encoder = self.encoders[i]
            if mode == "fit":
                graph, embed = encoder.init(inputs[idx], mode="fit")
            elif mode == "transform":
                graph, embed = encoder.init(self.data[i], mode="transform", query=inputs[idx],
                                            ref_data=self.graphs[i], ref_embeds=self.embeds[i])
            else:
                # neighbours in embedding space, initial value = weighted mean of the data rows
                graph, embed = encoder.init(self.embeds[i].detach(), mode="invert", query=inputs[idx],
                                            ref_data=self.graphs[i], ref_embeds=self.data[i], a=self.a, b=self.b)
            graphs.append(graph)
            embeds.append(embed)
        return graphs, embeds

    def _init_fit_distributed(self, inputs, encoder_indices):
        """Multi-GPU form of the fit initialisation: all graphs first (row-sharded kNN + replicated
        sigma/union), then every rank solves the spectral problems it owns (modality m -> rank m mod W)
        concurrently, then the results are broadcast -- instead of serialising the solves behind each
        other's broadcasts."""
        graphs, syms = [], []
        for idx, i in enumerate(encoder_indices):
            enc = self.encoders[i]
            graph = enc.fuzzy_knn_graph(inputs[idx], "fit", None, None, num_iters=10)
            g = _as_graph(graph)
            with profiler.stage("fuzzy_union"):
                sym = G.fuzzy_union(g.col2d, g.w2d)                      # model.py:271
            syms.append(sym)
            graphs.append(_coo(sym))
        embeds = [None] * len(encoder_indices)
        from umap_b200 import spectral as SP
        with profiler.stage("spectral_init"):
            # large graphs: one collective solve with the operator applications sharded by rows (identical result on
            # every rank, nothing to broadcast); small graphs: modality m on rank m mod W, concurrently, then broadcast
            collective = [SP.shardable(syms[idx], self.out_dim) for idx in range(len(encoder_indices))]
            for idx, i in enumerate(encoder_indices):
                if collective[idx]:
                    embeds[idx] = spectral_init(syms[idx], self.out_dim, shard=True).contiguous()
                elif D.rank() == self.encoders[i].id % D.world():
                    embeds[idx] = spectral_init(syms[idx], self.out_dim).contiguous()
            for idx, i in enumerate(encoder_indices):
                if collective[idx]:
                    continue
                owner = self.encoders[i].id % D.world()
                if embeds[idx] is None:
                    embeds[idx] = torch.empty((syms[idx].n_rows, self.out_dim), dtype=torch.float32, device="cuda")
                D.broadcast(embeds[idx], owner)
        return graphs, embeds

    # ------------------------------------------------------------------ checkpoint
    def save_state_dict(self, path: str) -> None:
        """ref: model.py:653-683 (same keys, torch-only payload)."""
        print("Warning: save_state_dict() saves the entire model state, which includes the source dataset. "
              "Make sure this is intended before proceeding.")
        state_dict = {
            "k_neighbors": self.k_neighbors,
            "out_dim": self.out_dim,
            "min_dist": self.min_dist,
            "num_encoders": self.num_encoders,
            "a": self.a,
            "b": self.b,
            "encoders": [{"sigmas": e.sigmas, "rhos": e.rhos} for e in self.encoders],
            "data": self.data,
            "graphs": self.graphs,
            "embeds": self.embeds,
        }
        dirname = os.path.dirname(path)
        if dirname and not os.path.exists(dirname):
            os.makedirs(dirname)
        torch.save(state_dict, path)

    @classmethod
    def load_state_dict(cls, path: str) -> "UMAPMixture":
        """ref: model.py:685-714."""
        state_dict = torch.load(path)
        model = cls.__new__(cls)
        model.k_neighbors = state_dict["k_neighbors"]
        model.out_dim = state_dict["out_dim"]
        model.min_dist = state_dict["min_dist"]
        model.num_encoders = state_dict["num_encoders"]
        model.a = state_dict["a"]
        model.b = state_dict["b"]
        model.encoders = [UMAPEncoder(model.k_neighbors, model.out_dim, id=i) for i in range(model.num_encoders)]
        for encoder, encoder_state in zip(model.encoders, state_dict["encoders"]):
            encoder.sigmas = encoder_state["sigmas"]
            encoder.rhos = encoder_state["rhos"]
        model.data = state_dict["data"]
        model.graphs = state_dict["graphs"]
        model.embeds = state_dict["embeds"]
        model._engine_defaults()
        return model
