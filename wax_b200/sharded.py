"""Row-sharded vector search across GPUs: one process (rank) per GPU, one CUDA engine per rank.

Design (BASELINE.json north_star): contiguous row ranges per rank, the query
replicated, the identical fused kernel on every shard, ONE exchange of the per-shard top-k candidates
(k x 24 B per rank) and a merge under the same total order (distance ascending, GLOBAL row ascending), so
results do not depend on the shard count.  Nothing else is exchanged.

Two transports for that exchange:
  * "p2p-fused" (default whenever the ranks can map each other's memory): the exchange is part of the scan launch --
    the kernel's last CTA writes its candidates into every rank's mailbox over NVLink/NVSwitch peer memory, waits for
    the others' flags and merges on the device, delivering the result into mapped host memory (wax_vs_shard_search,
    wax_b200/csrc/waxvs_shard.cuh).  torch.distributed is used ONCE, at construction, to pass the 128-byte mailbox
    handles around.  No collective launch, no D2H copy, no host merge per query.
  * "allgather": torch.distributed all_gather_into_tensor (NCCL) + host merge -- k > 128, the micro-batched /
    batched forms, and the CPU (gloo) tests of the host logic.

The reference has no distributed code at all ("none exist"); this is the only
parallelism the build adds.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

CAND_DTYPE = np.dtype([("distance", "<f4"), ("valid", "<u4"), ("row", "<u8"), ("frame_id", "<u8")])
assert CAND_DTYPE.itemsize == 24


def shard_range(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi): rows r*N/R .. (r+1)*N/R (global row = lo + local row)."""
    lo = (total_rows * rank) // world_size
    hi = (total_rows * (rank + 1)) // world_size
    return lo, hi


def clamp_topk(top_k: int) -> int:
    """clampTopK (MetalVectorEngine.swift:842-846)."""
    return 1 if top_k < 1 else min(int(top_k), 10_000)


def merge_candidates(gathered: np.ndarray, top_k: int) -> np.ndarray:
    """Host-side R-way merge: `gathered` is any array of CAND_DTYPE records (all ranks' lists);
    returns the best `top_k` valid ones ordered by (distance, global row)."""
    flat = gathered.reshape(-1)
    flat = flat[flat["valid"] != 0]
    if flat.size == 0:
        return flat
    order = np.lexsort((flat["row"], flat["distance"]))  # primary: distance, secondary: row
    return flat[order[:top_k]]


def merge_candidates_batch(gathered: np.ndarray, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised R-way merge for a batch: `gathered` is [world, batch, k] CAND_DTYPE; returns ([batch, top_k]
    records ordered by (distance, global row) with the invalid ones last, [batch] count of valid results).
    Same total order as merge_candidates (two stable sorts: by row, then by distance)."""
    world, batch, k = gathered.shape
    flat = np.ascontiguousarray(gathered.transpose(1, 0, 2)).reshape(batch, world * k)
    dist = np.where(flat["valid"] != 0, flat["distance"], np.float32(np.inf))
    by_row = np.argsort(flat["row"], axis=1, kind="stable")
    by_dist = np.argsort(np.take_along_axis(dist, by_row, 1), axis=1, kind="stable")
    order = np.take_along_axis(by_row, by_dist, 1)[:, :top_k]
    best = np.take_along_axis(flat, order, 1)
    return best, (best["valid"] != 0).sum(axis=1).astype(np.uint32)


def score_from_distance(similarity: int, d: np.ndarray) -> np.ndarray:
    """VectorMetric.score(fromDistance:) (VectorMetric.swift:32-43), vectorised, fp32."""
    d = d.astype(np.float32)
    s = (np.float32(1) - d) if similarity == 0 else -d
    return np.where(np.isfinite(d), s, np.float32(0)).astype(np.float32)


class ShardedVectorEngine:
    """`VectorSearchEngine.search` over a corpus row-sharded across the ranks of a torch.distributed group.

    local_search: optional injection point used by the CPU (gloo) tests of the host-side logic -- a callable
    (query ndarray, k) -> ndarray[CAND_DTYPE] of length k.  In production it is None and the local step is
    wax_vs_search_device on the rank's GPU.
    """

    def __init__(self, metric, dimensions: int, total_rows: int = 0, group=None,
                 local_search: Optional[Callable[[np.ndarray, int], np.ndarray]] = None, device=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.metric = metric
        self.dimensions = int(dimensions)
        self.total_rows = int(total_rows)
        self.row_lo, self.row_hi = shard_range(self.total_rows, self.world_size, self.rank)
        self._local_search = local_search
        self.engine = None
        self._bufs = {}
        self._comm_stream = None
        if local_search is None:
            from .engine import CUDAVectorEngine
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
            self.engine = CUDAVectorEngine(metric, dimensions, device=self.device.index)
            self._comm_stream = torch.cuda.Stream(device=self.device)   # all-gather + D2H overlap the next scan
            # two scan streams: consecutive queries alternate, so one scan's tail overlaps the next one's prologue
            self._scan_streams = [torch.cuda.Stream(device=self.device) for _ in range(2)]
            # Overlapping two scans pays when a shard is small (kernel tails/prologues overlap: +8 % at 7.7 GB,
            # +15 % at 1.9 GB) and costs ~3 % when it is large (19 GB: twice the bytes in flight per SM), measured
            # in profiles/bench_r01_n*_m*.json -- so alternate streams only below 12 GB per shard.
            self._overlap_scans = (self.row_hi - self.row_lo) * self.dimensions * 4 < 12e9
            self.transport = "allgather"
            self.transport_note = ""
            self._connect_peers()
        else:
            self.device = torch.device("cpu")
            self.transport = "allgather"

    def _connect_peers(self) -> None:
        """Create this rank's mailbox, exchange the handles (one all-gather of 128 bytes per rank, the only use of
        torch.distributed on this path) and map the peers' mailboxes.  All ranks agree on the outcome: if any rank
        cannot map a peer (no P2P / IPC path) every rank stays on the all-gather transport."""
        torch, dist = self._torch, self._dist
        from . import _lib as L
        from .engine import WaxError
        blob, ok = bytes(L.SHARD_HANDLE_BYTES), self.world_size <= L.SHARD_MAX_RANKS
        if ok:
            try:
                blob = self.engine.shard_open(self.rank, self.world_size, self.row_lo)
            except WaxError as exc:
                ok, self.transport_note = False, str(exc)
        if self.world_size > 1:
            dev = self.device if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
            mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
            every = torch.empty(self.world_size * L.SHARD_HANDLE_BYTES, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(every, mine, group=self.group)
            flat = every.cpu().numpy().tobytes()
            if ok:
                try:
                    self.engine.shard_connect([flat[i * L.SHARD_HANDLE_BYTES:(i + 1) * L.SHARD_HANDLE_BYTES]
                                               for i in range(self.world_size)])
                except WaxError as exc:
                    ok, self.transport_note = False, str(exc)
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if ok and flag.item() != 1:
                ok, self.transport_note = False, "another rank could not map its peers"
        if ok:
            self.transport = "p2p-fused"
        else:
            self.engine.shard_close()

    def close(self) -> None:
        """Collective: unmap the peers' mailboxes, agree that everyone has (barrier), then release the engine."""
        if self.engine is None or getattr(self, "_closed", False):
            return
        self._closed = True
        if self.transport == "p2p-fused":
            self.engine.shard_close()
            self.transport = "allgather"
        if self.world_size > 1 and self._dist.is_initialized():
            self._dist.barrier(group=self.group)
        self.engine.close()

    # -- corpus
    def fill_synthetic(self, seed: int, normalize: bool = True) -> None:
        """Each rank generates its own shard on device; frameId = global row."""
        self.engine.fill_synthetic(seed, self.row_hi - self.row_lo, first_row=self.row_lo, id_base=self.row_lo,
                                   normalize=normalize)

    # -- search
    def _buffers(self, k: int, slot: int = 0):
        torch = self._torch
        key = (k, slot)
        if key not in self._bufs:
            local = torch.zeros(k * 24, dtype=torch.uint8, device=self.device)
            gathered = torch.zeros(self.world_size * k * 24, dtype=torch.uint8, device=self.device)
            host = torch.zeros(self.world_size * k * 24, dtype=torch.uint8,
                               pin_memory=(self.device.type == "cuda"))
            self._bufs[key] = (local, gathered, host)
        return self._bufs[key]

    def search_async(self, d_query, top_k: int, slot: int = 0):
        """Enqueue local scan + all-gather + D2H on the current stream; returns a handle for finish().
        Queries are independent, so several may be in flight: give each a distinct `slot` (its result buffers)
        and finish() them in order -- the host merge of query i then overlaps the scan of query i+1."""
        torch, dist = self._torch, self._dist
        k = clamp_topk(top_k)
        local, gathered, host = self._buffers(k, slot)
        if self._local_search is not None:
            cands = np.ascontiguousarray(self._local_search(np.asarray(d_query, np.float32), k), dtype=CAND_DTYPE)
            local.copy_(torch.from_numpy(cands.view(np.uint8).reshape(-1).copy()))
        else:
            from . import _lib as L
            stream = torch.cuda.current_stream(self.device)
            rc = L.lib().wax_vs_search_device(self.engine.handle, C.c_void_p(d_query.data_ptr()), 1, k,
                                              self.row_lo, C.c_void_p(local.data_ptr()),
                                              C.c_void_p(stream.cuda_stream))
            if rc != 0:
                raise RuntimeError(f"wax_vs_search_device rc={rc}: {L.last_error()}")
        if self._comm_stream is not None:
            # exchange + D2H on the communication stream: the compute stream is free to start the next query
            scanned = torch.cuda.Event()
            scanned.record()
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(scanned)
                if self.world_size > 1:
                    dist.all_gather_into_tensor(gathered, local, group=self.group)
                else:
                    gathered.copy_(local)
                host.copy_(gathered, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            return (host, ev, k)
        if self.world_size > 1:
            dist.all_gather_into_tensor(gathered, local, group=self.group)
        else:
            gathered.copy_(local)
        host.copy_(gathered)
        return (host, None, k)

    def search_many_async(self, d_queries, top_k: int, slot: int = 0):
        """G independent queries (`d_queries`: [G, dims] device tensor) with ONE exchange: G fused scans alternating
        over two streams, one all-gather of G*k candidates per rank, one D2H.  Returns a handle for finish_many()."""
        torch, dist = self._torch, self._dist
        from . import _lib as L
        g = int(d_queries.shape[0])
        k = clamp_topk(top_k)
        key = (k, g, slot)
        if key not in self._bufs:
            local = torch.zeros(g * k * 24, dtype=torch.uint8, device=self.device)
            gathered = torch.zeros(self.world_size * g * k * 24, dtype=torch.uint8, device=self.device)
            host = torch.zeros(self.world_size * g * k * 24, dtype=torch.uint8, pin_memory=True)
            self._bufs[key] = (local, gathered, host)
        local, gathered, host = self._bufs[key]
        ready = torch.cuda.Event()
        ready.record()                                   # queries were produced on the current stream
        scanned = []
        n_streams = min(2 if self._overlap_scans else 1, g)
        for st in self._scan_streams[:n_streams]:
            st.wait_event(ready)
        q_base, q_stride = d_queries.data_ptr(), d_queries.stride(0) * 4
        for i in range(g):
            st = self._scan_streams[i % n_streams]
            rc = L.lib().wax_vs_search_device(self.engine.handle, C.c_void_p(q_base + i * q_stride), 1, k,
                                              self.row_lo, C.c_void_p(local.data_ptr() + i * k * 24),
                                              C.c_void_p(st.cuda_stream))
            if rc != 0:
                raise RuntimeError(f"wax_vs_search_device rc={rc}: {L.last_error()}")
        for st in self._scan_streams[:n_streams]:
            ev = torch.cuda.Event()
            ev.record(st)
            scanned.append(ev)
        with torch.cuda.stream(self._comm_stream):
            for ev in scanned:
                self._comm_stream.wait_event(ev)
            if self.world_size > 1:
                dist.all_gather_into_tensor(gathered, local, group=self.group)
            else:
                gathered.copy_(local)
            host.copy_(gathered, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        return (host, done, k, g)

    def finish_many(self, handle) -> List[List[Tuple[int, float]]]:
        host, done, k, g = handle
        done.synchronize()
        cands = host.numpy().view(CAND_DTYPE).reshape(self.world_size, g, k)
        k_eff = min(k, self.total_rows) if self.total_rows else k
        sim = self.metric.to_vec_similarity()
        out = []
        for i in range(g):
            best = merge_candidates(cands[:, i, :], k_eff)
            scores = score_from_distance(sim, best["distance"])
            out.append([(int(best["frame_id"][j]), float(scores[j])) for j in range(best.size)])
        return out

    def _merge_on_device(self, gathered, b: int, k: int, k_eff: int, stream):
        """[world][b][k] gathered candidates (device bytes) -> [b][k_eff] merged candidates on the device, by the library's
        merge kernel (same (distance, GLOBAL row) rule as the fused exchange); enqueued on `stream`."""
        from . import _lib as L
        merged = self._torch.empty(b * k_eff * 24, dtype=self._torch.uint8, device=self.device)
        rc = L.lib().wax_vs_merge_candidates_device(self.engine.handle, C.c_void_p(gathered.data_ptr()), self.world_size, b, k,
                                                    k_eff, C.c_void_p(merged.data_ptr()), C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"wax_vs_merge_candidates_device rc={rc}: {L.last_error()}")
        return merged

    def _unpack_merged(self, host_bytes: np.ndarray, b: int, k_eff: int):
        best = host_bytes.view(CAND_DTYPE).reshape(b, k_eff)
        scores = score_from_distance(self.metric.to_vec_similarity(), best["distance"])
        return best["frame_id"].astype(np.uint64), scores, (best["valid"] != 0).sum(axis=1).astype(np.uint32)

    def search_batch_arrays(self, queries, top_k: int):
        """A batch of independent queries against the sharded corpus: every rank runs the batched tensor-core levels
        (wax_vs_search_batch_device: bf16-shadow nominations -> TF32 retry -> exact scan, results identical to
        single-query scans) on its shard, ONE all-gather carries batch x k candidates per rank, ONE merge kernel
        (wax_vs_merge_candidates_device) ranks them on the device.  `queries`: [batch, dims] host array or device tensor (identical on every rank).  Returns
        (ids [batch, k_eff] uint64, scores [batch, k_eff] float32, n_valid [batch] uint32)."""
        torch, dist = self._torch, self._dist
        k = clamp_topk(top_k)
        k_eff = min(k, self.total_rows) if self.total_rows else k
        if self._local_search is not None:
            qs = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dimensions)
            b = qs.shape[0]
            local_np = np.zeros((b, k), CAND_DTYPE)
            for i in range(b):
                local_np[i] = np.ascontiguousarray(self._local_search(qs[i], k), dtype=CAND_DTYPE)
            local = torch.from_numpy(local_np.view(np.uint8).reshape(-1).copy())
        else:
            from . import _lib as L
            if isinstance(queries, torch.Tensor):
                d_qs = queries.to(self.device, dtype=torch.float32).contiguous().reshape(-1, self.dimensions)
            else:
                d_qs = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dimensions)).to(self.device)
            b = int(d_qs.shape[0])
            local = torch.empty(b * k * 24, dtype=torch.uint8, device=self.device)
            stream = torch.cuda.current_stream(self.device)
            rc = L.lib().wax_vs_search_batch_device(self.engine.handle, C.c_void_p(d_qs.data_ptr()), b, k, self.row_lo,
                                                    C.c_void_p(local.data_ptr()), C.c_void_p(stream.cuda_stream))
            if rc != 0:
                raise RuntimeError(f"wax_vs_search_batch_device rc={rc}: {L.last_error()}")
        if b == 0 or self.total_rows == 0:
            return np.zeros((b, 0), np.uint64), np.zeros((b, 0), np.float32), np.zeros(b, np.uint32)
        if self.world_size > 1:
            gathered = torch.empty(self.world_size * b * k * 24, dtype=torch.uint8, device=local.device)
            dist.all_gather_into_tensor(gathered, local, group=self.group)
        else:
            gathered = local
        if self._local_search is None:      # GPU: merge on the device, one D2H of the final batch x k_eff records
            merged = self._merge_on_device(gathered, b, k, k_eff, torch.cuda.current_stream(self.device))
            return self._unpack_merged(merged.cpu().numpy(), b, k_eff)
        cands = gathered.cpu().numpy().view(CAND_DTYPE).reshape(self.world_size, b, k)
        best, n_valid = merge_candidates_batch(cands, k_eff)
        scores = score_from_distance(self.metric.to_vec_similarity(), best["distance"])
        return best["frame_id"].astype(np.uint64), scores, n_valid

    # -- pipelined batches: the host merge of batch i overlaps the tensor-core pass of batch i+1 ----------------------
    def search_batch_submit(self, queries, top_k: int):
        """Start a batch on the engine's worker thread (ONE worker: batches, and therefore the all-gathers, are issued
        in submission order on every rank) and return a future for finish_batch().  The worker does the shard's
        tensor-core levels (a blocking C call that releases the GIL), the all-gather and the D2H copy; the caller is
        free to merge the previous batch meanwhile.  Queries must stay alive until finish_batch()."""
        torch, dist = self._torch, self._dist
        if getattr(self, "_worker", None) is None:
            from concurrent.futures import ThreadPoolExecutor
            self._worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="waxvs-shard")
            self._batch_stream = torch.cuda.Stream(device=self.device) if self._local_search is None else None
        k = clamp_topk(top_k)
        if self._local_search is not None:
            # CPU (gloo) form used by the tests of the host logic: the injected local search stands in for the GPU
            # step; the worker thread, the submission-ordered all-gathers and the merge are the production ones.
            qs = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dimensions).copy()
            b = qs.shape[0]

            def work_cpu():
                local_np = np.zeros((b, k), CAND_DTYPE)
                for i in range(b):
                    local_np[i] = np.ascontiguousarray(self._local_search(qs[i], k), dtype=CAND_DTYPE)
                local = torch.from_numpy(local_np.view(np.uint8).reshape(-1).copy())
                if self.world_size > 1:
                    gathered = torch.empty(self.world_size * b * k * 24, dtype=torch.uint8)
                    dist.all_gather_into_tensor(gathered, local, group=self.group)
                else:
                    gathered = local
                return gathered, None, None, None

            return (self._worker.submit(work_cpu), b, k)
        if isinstance(queries, torch.Tensor):
            d_qs = queries.to(self.device, dtype=torch.float32).contiguous().reshape(-1, self.dimensions)
        else:
            d_qs = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dimensions)).to(self.device)
        ready = torch.cuda.Event()
        ready.record()                                   # the queries were produced on the caller's stream
        b = int(d_qs.shape[0])

        def work():
            from . import _lib as L
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(self._batch_stream):
                self._batch_stream.wait_event(ready)
                local = torch.empty(b * k * 24, dtype=torch.uint8, device=self.device)
                rc = L.lib().wax_vs_search_batch_device(self.engine.handle, C.c_void_p(d_qs.data_ptr()), b, k, self.row_lo,
                                                        C.c_void_p(local.data_ptr()), C.c_void_p(self._batch_stream.cuda_stream))
                if rc != 0:
                    raise RuntimeError(f"wax_vs_search_batch_device rc={rc}: {L.last_error()}")
                if self.world_size > 1:
                    gathered = torch.empty(self.world_size * b * k * 24, dtype=torch.uint8, device=self.device)
                    dist.all_gather_into_tensor(gathered, local, group=self.group)
                else:
                    gathered = local
                k_eff = min(k, self.total_rows) if self.total_rows else k
                merged = self._merge_on_device(gathered, b, k, k_eff, self._batch_stream) if b and self.total_rows else gathered
                host = torch.empty(merged.numel(), dtype=torch.uint8, pin_memory=True)
                host.copy_(merged, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
            return host, done, (gathered, merged), d_qs  # keep the device buffers alive until the copy has finished

        return (self._worker.submit(work), b, k)

    def finish_batch(self, handle):
        """Wait for a submitted batch and merge it: (ids [batch, k_eff], scores, n_valid) as search_batch_arrays."""
        fut, b, k = handle
        host, done, _gathered, _d_qs = fut.result()
        if done is not None:
            done.synchronize()
        k_eff = min(k, self.total_rows) if self.total_rows else k
        if b == 0 or self.total_rows == 0:
            return np.zeros((b, 0), np.uint64), np.zeros((b, 0), np.float32), np.zeros(b, np.uint32)
        if self._local_search is None:      # GPU: the worker already merged on the device
            return self._unpack_merged(host.numpy(), b, k_eff)
        cands = host.numpy().view(CAND_DTYPE).reshape(self.world_size, b, k)
        best, n_valid = merge_candidates_batch(cands, k_eff)
        scores = score_from_distance(self.metric.to_vec_similarity(), best["distance"])
        return best["frame_id"].astype(np.uint64), scores, n_valid

    def search_batch(self, queries, top_k: int) -> List[List[Tuple[int, float]]]:
        ids, scores, ns = self.search_batch_arrays(queries, top_k)
        return [[(int(ids[i, j]), float(scores[i, j])) for j in range(int(ns[i]))] for i in range(ids.shape[0])]

    def finish(self, handle) -> List[Tuple[int, float]]:
        host, ev, k = handle
        if ev is not None:
            ev.synchronize()
        cands = host.numpy().view(CAND_DTYPE)
        k_eff = min(k, self.total_rows) if self.total_rows else k
        best = merge_candidates(cands, k_eff)
        scores = score_from_distance(self.metric.to_vec_similarity(), best["distance"])
        return [(int(best["frame_id"][i]), float(scores[i])) for i in range(best.size)]

    def search(self, vector: Sequence[float], top_k: int) -> List[Tuple[int, float]]:
        torch = self._torch
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        if q.size != self.dimensions:
            from .engine import EncodingError
            raise EncodingError(f"vector dimension mismatch: expected {self.dimensions}, got {q.size}")
        if self.total_rows == 0:
            return []
        if self._local_search is not None:
            return self.finish(self.search_async(q, top_k))
        if self.transport == "p2p-fused" and clamp_topk(top_k) <= 128:
            return self._search_fused(q, top_k)
        d_q = torch.from_numpy(q).to(self.device, non_blocking=False)
        return self.finish(self.search_async(d_q, top_k))

    def search_filtered(self, vector: Sequence[float], top_k: int, allow=None, deny=None) -> List[Tuple[int, float]]:
        """Filtered search over the whole sharded corpus (collective: same query and ids on every rank).  Needs the
        fused peer-memory transport: the filter rides in each rank's scan, the exchange is unchanged."""
        q = np.ascontiguousarray(vector, dtype=np.float32).reshape(-1)
        if q.size != self.dimensions:
            from .engine import EncodingError
            raise EncodingError(f"vector dimension mismatch: expected {self.dimensions}, got {q.size}")
        if self.total_rows == 0:
            return []
        if self.transport != "p2p-fused" or clamp_topk(top_k) > 128:
            from .engine import InvalidToc
            raise InvalidToc("sharded filtered search needs the p2p-fused transport and top_k <= 128")
        return self.engine.shard_search_filtered(q, top_k, allow=allow, deny=deny)

    def _search_fused(self, q: np.ndarray, top_k: int) -> List[Tuple[int, float]]:
        """wax_vs_shard_search: host query in, merged host result out; scan + NVLink exchange + merge in one launch."""
        return self.engine.shard_search(q, top_k)

    def time_search(self, top_k: int, iters: int, warmup: int = 3, n_queries: int = 1, seed: int = 7):
        """Device-timed collective searches, strictly one at a time on one stream (the sharded twin of
        CUDAVectorEngine.time_search).  Returns (ms_total, kernel launches)."""
        return self.engine.time_shard_search(top_k, iters, warmup=warmup, n_queries=n_queries, seed=seed)
