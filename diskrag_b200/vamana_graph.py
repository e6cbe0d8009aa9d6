"""Drop-in for pydiskann/vamana_graph.py: the same public names and call signatures, array-backed and
executed on the GPU through the C ABI (no CPU fallback).

Reference surface mirrored here:
  Node, VamanaGraphWithPQ, VamanaGraph                         vamana_graph.py:8-56, 234-245
  build_vamana, build_vamana_with_pq                           :464-533, 686-688
  greedy_search, greedy_search_optimized                       :607-640, 762-793   (variant B)
  greedy_search_with_pq                                        :357-400            (NameError in the reference; works here)
  beam_search_from_disk                                        :719-760            (variant D)
  beam_search, beam_search_with_pq                             :535-605, 690-717   (variant C, see note)
  compute_distance, compute_query_distance, l2_distance_fast,
  pq_distance_fast, compute_approximate_medoid,
  robust_prune, robust_prune_with_pq                           :249-355, 402-450, 642-683
Additive: `search_batch(graph_or_reader, Q, ...)` — the batched entry point the reference lacks.

The graph keeps `vec f32[N,D]`, `adj u32[N,w]` + `deg` (w >= R: like the reference's neighbour sets, a row may
outgrow R through the reverse edges of insert_node, :106-114), `codes u8[N,M]`, a lazy-delete mask, and mirrors them
into one device index (GpuIndex).  The mirror is patched in place after a mutation — appended nodes, changed rows,
changed delete flags (dr_index_append / dr_index_patch_rows / ...): O(rows touched), not O(N*D) — and rebuilt only
when the row width or the PQ model changes.  On the device an unused slot of a row holds 0xFFFFFFFF = "no neighbour":
the reference's in-memory searches iterate `node.neighbors` (true degree, the set's iteration order, no padding;
:607-640, cython_utils.pyx:108).  The 0-padding that IS a neighbour belongs to the on-disk path only (index.dat rows
read back by MMapNodeReader: `to_records`, `beam_search_from_disk`, `GpuIndex.from_records`).
`graph.nodes` is a lazy mapping that materialises `Node` objects on demand, so a 1M-node graph does not cost GBs of
Python objects; a materialised Node keeps its own `set`, so the iteration order the reference would see is the order
the device row is written in.

Variant C note: the reference's beam_search_with_pq caps the result heap at k and truncates its frontier
by popping the *best* entries (:592-593), which yields recall 0.036 on its own benchmark shape (SURVEY
§3.3).  That is a defect, not a contract; here beam_search[_with_pq] returns the k best of a correct search
with list size max(k, beam_width), as (sqrt(d), id) tuples like the reference.
"""
import numpy as np

from . import ops
from ._lib import as_f32, check, lib, ptr
from .cython_utils import (build_vamana_index_cython, compute_approximate_medoid_cython, cosine_similarity_cython,  # noqa: F401
                           generate_initial_neighbors_cython, greedy_search_cython, l2_distance_fast_cython,
                           pq_distance_fast_cython, robust_prune_cython)
from .engine import GpuIndex


class Node:
    def __init__(self, idx, vector, pq_code=None, is_deleted=False):
        self.idx = idx
        self.vector = vector
        self.pq_code = pq_code
        self.neighbors = set()
        self.is_deleted = is_deleted


class _Nodes:
    """dict-like view {idx: Node} over the graph arrays; Node objects are created on first access and written
    back (neighbours, deletion flag) before the arrays are used again."""

    def __init__(self, graph):
        self._g = graph
        self._live = {}

    def __len__(self):
        return self._g._n

    def __contains__(self, idx):
        return isinstance(idx, (int, np.integer)) and 0 <= idx < self._g._n

    def __iter__(self):
        return iter(range(self._g._n))

    def keys(self):
        return range(self._g._n)

    def values(self):
        return (self[i] for i in range(self._g._n))

    def items(self):
        return ((i, self[i]) for i in range(self._g._n))

    def __getitem__(self, idx):
        idx = int(idx)
        node = self._live.get(idx)
        if node is None:
            g = self._g
            if not 0 <= idx < g._n:
                raise KeyError(idx)
            node = Node(idx, g._vec[idx], g._codes[idx] if g._codes is not None else None, bool(g._deleted[idx]))
            node.neighbors = set(int(x) for x in g._adj[idx, :g._deg[idx]])
            node._snapshot = (frozenset(node.neighbors), node.is_deleted)
            self._live[idx] = node
        return node

    def __setitem__(self, idx, node):
        g = self._g
        idx = int(idx)
        if idx == g._n:
            g._append(node.vector, node.pq_code)
        elif not 0 <= idx < g._n:
            raise KeyError(f"node ids must be dense: next id is {g._n}, got {idx}")
        node._snapshot = None
        self._live[idx] = node

    def _flush(self):
        """Write materialised nodes that changed back into the arrays (rows in the set's iteration order, never truncated:
        the arrays widen instead) and note which rows / flags the device mirror has to patch."""
        g = self._g
        changed = False
        for idx, node in self._live.items():
            snap = getattr(node, "_snapshot", None)
            cur = (frozenset(node.neighbors), bool(node.is_deleted))
            if snap == cur:
                continue
            nb = list(node.neighbors)
            if len(nb) > g._adj.shape[1]:
                g._widen(len(nb))
            if snap is None or snap[0] != cur[0]:
                g._adj[idx, :len(nb)] = nb
                g._adj[idx, len(nb):] = 0
                g._deg[idx] = len(nb)
                g._rows_dirty.add(idx)
            if snap is None or snap[1] != cur[1]:
                g._deleted[idx] = node.is_deleted
                g._del_dirty.add(idx)
            if node.pq_code is not None and g._codes is not None and not np.array_equal(g._codes[idx], node.pq_code):
                g._codes[idx] = node.pq_code
                g._vec_dirty.add(idx)
            node._snapshot = cur
            changed = True
        return changed


class VamanaGraphWithPQ:
    def __init__(self, R, pq_model=None, distance_metric='l2', device=0):
        self.R = R
        self.pq_model = pq_model
        self.medoid_idx = None
        self.use_pq_for_search = False
        self._distance_table_cache = {}
        self.distance_metric = distance_metric
        self.device = device
        self._n = 0
        self._vec = None
        self._adj = np.zeros((0, R), np.uint32)
        self._deg = np.zeros(0, np.int32)
        self._codes = None
        self._deleted = np.zeros(0, bool)
        self._dirty = True            # the device mirror must be rebuilt from scratch (row width / PQ model / bulk change)
        self._rows_dirty = set()      # rows whose adjacency changed since the mirror was last patched
        self._vec_dirty = set()       # rows whose vector / code changed (a deleted id re-enabled with a new vector)
        self._del_dirty = set()       # rows whose delete flag changed
        self._gpu = None
        self._gpu_n = 0               # nodes the mirror holds
        self._vec_store = None; self._codes_store = None
        self._adj_store, self._deg_store, self._del_store = self._adj, self._deg, self._deleted
        self.nodes = _Nodes(self)

    # ---- array plumbing ------------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, vec, adj, deg=None, codes=None, pq_model=None, medoid_idx=0, R=None, distance_metric='l2', device=0):
        g = cls(R if R is not None else adj.shape[1], pq_model, distance_metric, device)
        g._vec = as_f32(vec)
        g._n = g._vec.shape[0]
        g._adj = np.ascontiguousarray(adj, np.uint32).copy()
        g._deg = (np.full(g._n, g._adj.shape[1], np.int32) if deg is None else np.ascontiguousarray(deg, np.int32).copy())
        g._codes = None if codes is None else np.ascontiguousarray(codes, np.uint8).copy()
        g._deleted = np.zeros(g._n, bool)
        g._adopt()
        g.medoid_idx = int(medoid_idx)
        return g

    def _reserve(self, n, D=None):
        """Host arrays grow geometrically (amortised O(1) per appended node); [:_n] is the live part."""
        cap = self._adj_store.shape[0]
        if self._vec_store is not None and n <= cap:
            return
        ncap = max(n, cap + cap // 2 + 64)

        def grow(store, tail, dtype):
            b = np.zeros((ncap,) + tail, dtype)
            if store is not None and self._n:
                b[:self._n] = store[:self._n]
            return b
        D = self._vec_store.shape[1] if self._vec_store is not None else D
        self._vec_store = grow(self._vec_store, (D,), np.float32)
        self._adj_store = grow(self._adj_store, (self._adj_store.shape[1],), np.uint32)
        self._deg_store = grow(self._deg_store, (), np.int32)
        self._del_store = grow(self._del_store, (), bool)
        if self._codes_store is not None:
            self._codes_store = grow(self._codes_store, (self._codes_store.shape[1],), np.uint8)
        self._views()

    def _views(self):
        n = self._n
        self._vec = self._vec_store[:n]; self._adj = self._adj_store[:n]; self._deg = self._deg_store[:n]
        self._deleted = self._del_store[:n]
        self._codes = None if self._codes_store is None else self._codes_store[:n]
        for idx, node in self.nodes._live.items():          # Node.vector stays a view of the (possibly moved) array
            if idx < n:
                node.vector = self._vec[idx]

    def _adopt(self):
        """bind the growable stores to freshly assigned arrays"""
        self._vec_store, self._adj_store, self._deg_store, self._del_store = self._vec, self._adj, self._deg, self._deleted
        self._codes_store = self._codes

    def _set_codes(self, codes):
        codes = np.ascontiguousarray(codes, np.uint8)
        self._codes_store = np.zeros((self._adj_store.shape[0], codes.shape[1]), np.uint8)
        self._codes_store[:self._n] = codes
        self._codes = self._codes_store[:self._n]

    def _widen(self, need):
        """A neighbour set outgrew the row width (reverse edges are never pruned, :106-114): widen every row."""
        w = max(need, self._adj_store.shape[1] * 2)
        a = np.zeros((self._adj_store.shape[0], w), np.uint32)
        a[:, :self._adj_store.shape[1]] = self._adj_store
        self._adj_store = a
        self._adj = a[:self._n]
        self._dirty = True                                   # the device row stride changes: rebuild the mirror

    def _append(self, vector, pq_code):
        v = as_f32(vector).reshape(-1)
        if pq_code is not None and self._codes_store is None:
            self._set_codes(np.zeros((self._n, np.asarray(pq_code).size), np.uint8))
        self._reserve(self._n + 1, v.size)
        i = self._n
        self._n += 1
        self._views()
        self._vec[i] = v
        self._adj[i] = 0; self._deg[i] = 0; self._deleted[i] = False
        if self._codes is not None:
            self._codes[i] = 0 if pq_code is None else np.asarray(pq_code, np.uint8).reshape(-1)

    def _sync(self):
        """Materialised nodes -> arrays; what changed is noted per row (_rows_dirty / _vec_dirty / _del_dirty) for gpu_index()."""
        self.nodes._flush()

    def to_records(self, R=None):
        """index.dat image, u32[N, D+R] (DiskANNPersist.save_index, diskann_persist.py:17-24): short rows 0-padded."""
        self._sync()
        R = self.R if R is None else R
        D = self._vec.shape[1]
        rec = np.zeros((self._n, D + R), np.uint32)
        rec[:, :D] = self._vec.view(np.uint32)
        w = min(R, self._adj.shape[1])
        adj = self._adj[:, :w].copy()
        adj[np.arange(w)[None, :] >= self._deg[:, None]] = 0
        rec[:, D:D + w] = adj
        return rec

    def _device_rows(self, rows=None):
        """Adjacency rows as the in-memory searches of the reference see them: true degree, unused slots = no neighbour."""
        adj = (self._adj if rows is None else self._adj[rows]).copy()
        deg = self._deg if rows is None else self._deg[rows]
        adj[np.arange(adj.shape[1])[None, :] >= deg[:, None]] = 0xFFFFFFFF
        return adj

    def gpu_index(self) -> GpuIndex:
        """Device mirror of the arrays, patched in place after mutations (O(rows touched))."""
        self._sync()
        has_pq = self.pq_model is not None and getattr(self.pq_model, "is_fitted", False) and self._codes is not None
        if self._gpu is not None and not self._dirty and (self._gpu.M > 0) != bool(has_pq):
            self._dirty = True
        if self._gpu is None or self._dirty:
            if self._gpu is not None:
                self._gpu.close()
            cb = None
            if has_pq:
                from .io.diskann_persist import codebook_of
                cb = self.pq_model.codebook() if hasattr(self.pq_model, "codebook") else codebook_of(self.pq_model)
            codes = self._codes if cb is not None else None
            self._gpu = GpuIndex.from_arrays(self._vec, self._device_rows(), codes, cb, self.medoid_idx or 0, self.device)
            if self._deleted.any():
                m = np.ascontiguousarray(self._deleted, np.uint8)
                check(lib().dr_index_set_deleted(self._gpu._h, ptr(m)), "dr_index_set_deleted")
            self._gpu_n = self._n
            self._dirty = False
            self._rows_dirty.clear(); self._vec_dirty.clear(); self._del_dirty.clear()
            return self._gpu
        g = self._gpu
        if self._n > self._gpu_n:                                              # appended nodes
            lo = self._gpu_n
            g.append(self._vec[lo:], self._codes[lo:] if g.M > 0 else None)
            self._rows_dirty.update(i for i in range(lo, self._n) if self._deg[i] > 0)
            self._del_dirty.update(i for i in range(lo, self._n) if self._deleted[i])
            self._gpu_n = self._n
        if self._vec_dirty:
            rows = np.fromiter(sorted(self._vec_dirty), np.int64)
            g.patch_vectors(rows, self._vec[rows], self._codes[rows] if g.M > 0 else None)
            self._vec_dirty.clear()
        if self._rows_dirty:
            rows = np.fromiter(sorted(self._rows_dirty), np.int64)
            g.patch_rows(rows, self._device_rows(rows))
            self._rows_dirty.clear()
        if self._del_dirty:
            rows = np.fromiter(sorted(self._del_dirty), np.int64)
            g.set_deleted_rows(rows, self._deleted[rows].astype(np.uint8))
            self._del_dirty.clear()
        return g

    # ---- reference API -----------------------------------------------------------------------------------
    def set_pq_model(self, pq_model):
        self.pq_model = pq_model
        if pq_model and pq_model.is_fitted and self._n > 0:
            print("重新編碼所有向量...")
            self._set_codes(pq_model.encode(self._vec))
            for idx, node in self.nodes._live.items():
                node.pq_code = self._codes[idx]
            self._dirty = True
            print(f"完成 {self._n} 個向量的 PQ 編碼")

    def add_node(self, idx, vector, pq_code=None):
        if pq_code is None and self.pq_model and self.pq_model.is_fitted:
            pq_code = self.pq_model.encode(np.asarray(vector).reshape(1, -1))[0]
        self.nodes[idx] = Node(idx, as_f32(vector), pq_code)

    def add_edge(self, from_idx, to_idx):
        if from_idx != to_idx and from_idx in self.nodes and to_idx in self.nodes:
            self.nodes[from_idx].neighbors.add(int(to_idx))

    def enable_pq_search(self, enable=True):
        if enable and (not self.pq_model or not self.pq_model.is_fitted):
            raise ValueError("PQ 模型未訓練，無法啟用 PQ 搜索")
        self.use_pq_for_search = enable
        print("已啟用 PQ 加速搜索" if enable else "已禁用 PQ 加速搜索，使用精確距離計算")

    def insert_node(self, idx, vector, pq_code=None, L_insert=None, *, robust=False):
        """vamana_graph.py:58-114: greedy search for L candidates from the medoid, `robust_prune_cython(alpha=1.0)`, reverse
        edges without a re-prune.  What the reference's prune does there is keep the R NEAREST candidates: its removal loop
        rebinds the list it is iterating over, so no candidate is ever dropped (cython_utils.pyx:147-165), and the exact
        distance is compute_distance's l2 branch (the metric string lands in the query_vector slot, vamana_graph.py:273).
        That is the default here — same neighbour sets as the reference; robust=True runs a real RobustPrune instead."""
        vector = as_f32(vector)
        if idx in self.nodes:
            node = self.nodes[idx]
            if node.is_deleted:
                node.is_deleted = False
                self._vec[idx] = vector
                node.vector = self._vec[idx]
                node.pq_code = pq_code
                node.neighbors.clear()
                if pq_code is not None and self._codes is not None:
                    self._codes[idx] = np.asarray(pq_code, np.uint8).reshape(-1)
                self._vec_dirty.add(int(idx))
                print(f"節點 {idx} 已重新啟用。")
            else:
                raise ValueError(f"節點 {idx} 已存在。")
        else:
            if pq_code is None and self.pq_model and self.pq_model.is_fitted:
                pq_code = self.pq_model.encode(vector.reshape(1, -1))[0]
            self.nodes[idx] = Node(idx, vector, pq_code)
        if len(self.nodes) == 1:
            self.medoid_idx = idx
            return
        start = self.medoid_idx
        if start is None or self.nodes[start].is_deleted:
            start = next((i for i in range(self._n) if not self._deleted[i] and i != idx), None)
            if start is None:
                self.medoid_idx = idx
                return
        L_val = L_insert if L_insert is not None else self.R * 2
        cands = greedy_search_cython(self, start, self.nodes[idx].vector, L_val, compute_query_distance)
        if robust:
            _graph_prune(self, idx, set(cands), 1.0, self.R)
        else:
            _graph_prune_nearest(self, idx, set(cands), self.R)
        for nb in list(self.nodes[idx].neighbors):
            if nb in self.nodes and not self.nodes[nb].is_deleted:
                self.add_edge(nb, idx)

    def delete_node(self, idx):
        if idx not in self.nodes:
            raise ValueError(f"節點 {idx} 不存在。")
        self.nodes[idx].is_deleted = True
        print(f"節點 {idx} 已標記為刪除。")

    def consolidate_index(self, R=None, L=None, alpha=None, pq_model=None, distance_metric=None, show_progress=False):
        """vamana_graph.py:127-230: rebuild from the live nodes (GPU build) and keep the original ids."""
        print("開始合併索引...")
        self._sync()
        live = np.flatnonzero(~self._deleted)
        if live.size == 0:
            print("沒有活動節點可供合併。索引已清空。")
            self.__init__(self.R, self.pq_model, self.distance_metric, self.device)
            return
        t = build_vamana_with_pq(self._vec[live], pq_model if pq_model is not None else self.pq_model,
                                 R if R is not None else self.R, L if L is not None else self.R * 2,
                                 alpha if alpha is not None else 1.0, False, show_progress,
                                 distance_metric if distance_metric is not None else self.distance_metric)
        # ids stay sparse in the reference (dict keyed by the old ids); the array form needs dense ids, so the
        # deleted slots stay in place as isolated, masked nodes
        self.nodes._flush()
        w = self._adj.shape[1]
        self._adj[:] = 0
        self._deg[:] = 0
        tw = min(w, t._adj.shape[1])
        mapped = live[t._adj[:, :tw].astype(np.int64)].astype(np.uint32)
        mapped[np.arange(tw)[None, :] >= t._deg[:, None]] = 0
        self._adj[live, :tw] = mapped
        self._deg[live] = np.minimum(t._deg, tw)
        if t._codes is not None:
            if self._codes is None:
                self._set_codes(np.zeros((self._n, t._codes.shape[1]), np.uint8))
            self._codes[live] = t._codes
        self.medoid_idx = int(live[t.medoid_idx]) if t.medoid_idx is not None else None
        self.nodes._live.clear()
        self._distance_table_cache.clear()
        self._dirty = True
        print(f"索引合併完成。剩餘 {live.size} 個活動節點。")

    def close(self):
        if self._gpu is not None:
            self._gpu.close()
            self._gpu = None


class VamanaGraph(VamanaGraphWithPQ):
    """Backward-compatible name (vamana_graph.py:234-245)."""

    def __init__(self, R):
        super().__init__(R)


# ---- scalar distance wrappers (vamana_graph.py:249-329) ---------------------------------------------------------
def l2_distance_fast(x, y):
    return l2_distance_fast_cython(x, y)


def pq_distance_fast(pq_model, code1, code2):
    if not pq_model or not pq_model.is_fitted:
        raise ValueError("PQ 模型未初始化")
    return pq_distance_fast_cython(pq_model, code1, code2)


def _pq_on(graph):
    return bool(getattr(graph, 'use_pq_for_search', False) and graph.pq_model and graph.pq_model.is_fitted)


def compute_distance(graph, idx1, idx2, query_vector=None, distance_metric='l2'):
    node1, node2 = graph.nodes[idx1], graph.nodes[idx2]
    if query_vector is not None and _pq_on(graph) and node2.pq_code is not None:
        table = graph.pq_model.compute_distance_table(query_vector)
        return graph.pq_model.asymmetric_distance_sq(node2.pq_code.reshape(1, -1), table)[0]
    if _pq_on(graph) and node1.pq_code is not None and node2.pq_code is not None:
        return pq_distance_fast(graph.pq_model, node1.pq_code, node2.pq_code)
    if distance_metric == 'l2':
        return l2_distance_fast(node1.vector, node2.vector)
    if distance_metric == 'cosine':
        return cosine_similarity_cython(node1.vector, node2.vector)
    raise ValueError(f"不支持的距離度量: {distance_metric}")


def compute_query_distance(graph, query_vector, node_idx, distance_metric='l2'):
    node = graph.nodes[node_idx]
    if _pq_on(graph) and node.pq_code is not None:
        qid = id(query_vector)
        if qid not in graph._distance_table_cache:
            graph._distance_table_cache[qid] = graph.pq_model.compute_distance_table(query_vector)
        return graph.pq_model.asymmetric_distance_sq(node.pq_code.reshape(1, -1), graph._distance_table_cache[qid])[0]
    if distance_metric == 'l2':
        return l2_distance_fast(node.vector, query_vector)
    if distance_metric == 'cosine':
        return cosine_similarity_cython(node.vector, query_vector)
    raise ValueError(f"不支持的距離度量: {distance_metric}")


def compute_approximate_medoid(points_array, sample_size=1000, batch_size=1024):
    pts = as_f32(points_array)
    n = pts.shape[0]
    samples = np.arange(n, dtype=np.int32) if n <= sample_size else \
        np.random.choice(n, sample_size, replace=False).astype(np.int32)
    return ops.medoid(pts, samples)


# ---- graph adapters ------------------------------------------------------------------------------------------------
def _as_graph(graph) -> VamanaGraphWithPQ:
    """Our graphs pass through; a foreign duck-typed graph (reference-style dict of Node objects) is converted
    once and cached on the object."""
    if isinstance(graph, VamanaGraphWithPQ):
        return graph
    g = getattr(graph, "_b200_mirror", None)
    if g is not None:
        # a foreign graph can be mutated behind the mirror's back (insert_node, delete_node, direct edits of node.neighbors): a cheap
        # fingerprint of what the searches depend on decides whether the mirror still stands
        n = len(graph.nodes)
        fp = (n, sum(len(graph.nodes[i].neighbors) for i in range(n)), sum(1 for i in range(n) if getattr(graph.nodes[i], "is_deleted", False)),
              getattr(graph, "medoid_idx", 0))
        if fp != getattr(g, "_foreign_fp", None):
            g.close()
            g = None
    if g is None:
        n = len(graph.nodes)
        R = max(int(getattr(graph, "R", 0)), max((len(graph.nodes[i].neighbors) for i in range(n)), default=0), 1)
        vec = np.stack([np.asarray(graph.nodes[i].vector, np.float32) for i in range(n)])
        adj = np.zeros((n, R), np.uint32); deg = np.zeros(n, np.int32)
        for i in range(n):
            nb = list(graph.nodes[i].neighbors)
            adj[i, :len(nb)] = nb; deg[i] = len(nb)
        codes = None
        if n and getattr(graph.nodes[0], "pq_code", None) is not None:
            codes = np.stack([graph.nodes[i].pq_code for i in range(n)]).astype(np.uint8)
        g = VamanaGraphWithPQ.from_arrays(vec, adj, deg, codes, getattr(graph, "pq_model", None),
                                          getattr(graph, "medoid_idx", 0) or 0, R, getattr(graph, "distance_metric", "l2"))
        g._deleted = np.array([bool(getattr(graph.nodes[i], "is_deleted", False)) for i in range(n)])
        g._adopt()
        g._foreign_fp = (n, int(deg.sum()), int(g._deleted.sum()), getattr(graph, "medoid_idx", 0))
        try:
            graph._b200_mirror = g
        except Exception:
            pass
    g.use_pq_for_search = bool(getattr(graph, "use_pq_for_search", False))
    return g


def _live_start(g, start_idx):
    """cython_utils.pyx:84-90: a deleted start node is replaced by the first live node."""
    g._sync()
    if g._deleted[start_idx]:
        live = np.flatnonzero(~g._deleted)
        if live.size == 0:
            return None
        return int(live[0])
    return int(start_idx)


def _graph_search(graph, start_idx, q, L, ignore_deleted=False):
    """greedy_search_cython semantics: <= L ids, ascending traversal distance.  ignore_deleted: the Python-level greedy_search /
    greedy_search_optimized never look at is_deleted (vamana_graph.py:607-640, 762-793)."""
    g = _as_graph(graph)
    if g.distance_metric not in ('l2', 'cosine') and not _pq_on(g):
        raise ValueError(f"unknown distance_metric {g.distance_metric!r}")
    start = int(start_idx) if ignore_deleted else _live_start(g, start_idx)
    if start is None:
        return []
    idx = g.gpu_index()
    # compute_query_distance (:301-329): ADC when PQ search is enabled, else 1 - cos for 'cosine', else squared L2
    dist = "pq" if _pq_on(g) else ("cosine" if g.distance_metric == 'cosine' else "exact")
    r = idx.search(q[None, :], k=1, L=L, W=1, dist=dist, rerank=False, want_list=True, start=start,
                   ignore_deleted=ignore_deleted)
    n = int(r.list_len[0])
    return [int(x) for x in r.list_ids[0, :n]]


def _prune_inputs(graph, point_idx, candidate_set, keep_self=False):
    g = _as_graph(graph)
    g._sync()
    cands = np.array(sorted(int(c) for c in candidate_set
                            if 0 <= int(c) < g._n and not g._deleted[int(c)] and (keep_self or int(c) != point_idx)), np.int64)
    return g, cands


def _store_neighbors(graph, g, point_idx, new):
    graph.nodes[point_idx].neighbors = new
    if g is not graph:
        g.nodes[point_idx].neighbors = set(new)


def _graph_prune(graph, point_idx, candidate_set, alpha, R):
    """A real RobustPrune (DiskANN alg. 2) on exact squared L2: what robust_prune_cython / robust_prune[_with_pq] are meant to
    compute.  The reference's own loops never drop a candidate (they rebind the list they iterate over, cython_utils.pyx:147-165,
    vamana_graph.py:424-441) and so return the R nearest; `_graph_prune_nearest` is that behaviour."""
    g, cands = _prune_inputs(graph, point_idx, candidate_set)
    sel = np.empty(max(R, 1), np.int32)
    import ctypes as C
    cv = np.ascontiguousarray(g._vec[cands]) if cands.size else np.zeros((0, g._vec.shape[1]), np.float32)
    pv = np.ascontiguousarray(g._vec[point_idx])
    cnt = C.c_int32(0)
    check(lib().dr_robust_prune(ptr(pv), ptr(cv), int(cands.size), g._vec.shape[1], float(alpha), int(R), ptr(sel), C.byref(cnt),
                                g.device), "dr_robust_prune")
    _store_neighbors(graph, g, point_idx, set(int(cands[i]) for i in sel[:cnt.value]))


def _graph_prune_nearest(graph, point_idx, candidate_set, R):
    """The reference's prune as it behaves (see _graph_prune): the R nearest live candidates by squared L2, ties by id
    (`candidates_with_dist.sort()` on (dist, cid) tuples), inserted into a fresh set in that order.
    Squared L2 also on a graph whose distance_metric is 'cosine': robust_prune_cython calls
    compute_distance_fn(graph, a, b, graph.distance_metric) (cython_utils.pyx:141,160), and compute_distance's fourth positional
    parameter is query_vector (vamana_graph.py:259) -- the metric never arrives, distance_metric keeps its default 'l2'.  (With
    use_pq_for_search on, that stray string reaches compute_distance_table and the reference's insert raises; here it prunes.)"""
    g, cands = _prune_inputs(graph, point_idx, candidate_set, keep_self=True)   # the reference does not exclude the point itself
    new = set()
    if cands.size:
        d = ops.l2sq_batch(np.ascontiguousarray(g._vec[cands]), np.ascontiguousarray(g._vec[point_idx])[None, :])
        for i in np.lexsort((cands, d))[:R]:
            new.add(int(cands[i]))
    _store_neighbors(graph, g, point_idx, new)


def robust_prune_with_pq(graph, point_idx, candidate_set, alpha, R, *, reference_semantics=False):
    (_graph_prune_nearest(graph, point_idx, candidate_set, R) if reference_semantics
     else _graph_prune(graph, point_idx, candidate_set, alpha, R))


def robust_prune(graph, point_idx, candidate_set, alpha, R, *, reference_semantics=False):
    robust_prune_with_pq(graph, point_idx, candidate_set, alpha, R, reference_semantics=reference_semantics)


def generate_initial_neighbors(n_points, R):
    return generate_initial_neighbors_cython(n_points, min(R, n_points - 1))


# ---- build (vamana_graph.py:464-533) -----------------------------------------------------------------------------
def build_vamana_with_pq(points, pq_model=None, R=16, L=32, alpha=1.2, use_pq_in_build=False, show_progress=False,
                         distance_metric='l2', seed=None, device=0):
    n_points = len(points)
    if n_points == 0:
        return VamanaGraphWithPQ(R, pq_model, distance_metric, device)
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.float32))
    codes = None
    if pq_model and pq_model.is_fitted:
        if show_progress:
            print("使用 PQ 編碼向量...")
        codes = pq_model.encode(pts)
    if show_progress:
        print("計算近似 medoid (GPU)...")
    medoid_idx = compute_approximate_medoid_cython(pts, sample_size=min(1000, n_points))
    if show_progress:
        print(f"選擇的近似 medoid: {medoid_idx}")
    if seed is None:
        import random
        seed = random.getrandbits(63)   # reproducible under random.seed(), like the reference's shuffles
    adj, deg = ops.vamana_build(pts, R, L, alpha, medoid_idx, seed, device)
    g = VamanaGraphWithPQ.from_arrays(pts, adj, deg, codes, pq_model, medoid_idx, R, distance_metric, device)
    g.use_pq_for_search = use_pq_in_build
    if show_progress:
        print(f"Vamana 圖構建完成，共 {n_points} 個節點")
        if pq_model:
            print(f"PQ 壓縮比: {pq_model.get_memory_usage()['compression_ratio']:.1f}x")
    return g


def build_vamana(points, R=16, L=32, alpha=1.2, show_progress=False):
    return build_vamana_with_pq(points, None, R, L, alpha, False, show_progress)


# ---- searches ------------------------------------------------------------------------------------------------------
def greedy_search(graph, start_idx, query_vector, L):
    """Variant B (:607-640).  With PQ enabled the reference dispatches to greedy_search_with_pq, which raises
    NameError (:387); here it runs the PQ traversal."""
    g = _as_graph(graph)
    return _graph_search(g, int(start_idx), as_f32(query_vector).ravel(), int(L), ignore_deleted=not _pq_on(g))


def greedy_search_optimized(graph, start_idx, query_vector, L):
    """:762-793: always exact, never looks at is_deleted."""
    g = _as_graph(graph)
    was, g.use_pq_for_search = g.use_pq_for_search, False
    try:
        return _graph_search(g, int(start_idx), as_f32(query_vector).ravel(), int(L), ignore_deleted=True)
    finally:
        g.use_pq_for_search = was


def greedy_search_with_pq(graph, start_idx, query_vector, L):
    """:357-400 (NameError in the reference): the PQ traversal, lazy deletes honoured."""
    return _graph_search(graph, int(start_idx), as_f32(query_vector).ravel(), int(L))


def beam_search_with_pq(graph, query_vector, start_idx=None, beam_width=5, k=3, use_pq=True, *, reference_semantics=False):
    """Variant C (:535-605).  By default the correct search (list size max(k, beam_width)): the reference's own loop throws
    the BEST frontier entries away when it truncates (:595-596), which costs it recall.  reference_semantics=True runs that
    loop as written (csrc/beam_c.cu, dr_beam_search_c) for callers that depend on the reference's exact answers; l2 only."""
    g = _as_graph(graph)
    if start_idx is None:
        start_idx = g.medoid_idx if g.medoid_idx is not None else 0
    start = _live_start(g, int(start_idx))
    if start is None:
        return []
    pq = bool(use_pq and g.pq_model and g.pq_model.is_fitted and g._codes is not None)
    idx = g.gpu_index()
    if reference_semantics:
        if not pq and getattr(g, "distance_metric", "l2") != "l2":
            raise ValueError("reference_semantics=True supports distance_metric='l2' only")
        r = idx.beam_search_c(as_f32(query_vector).reshape(1, -1), k=int(k), beam_width=int(beam_width),
                              dist="pq" if pq else "exact", sqrt_out=True, start=start)
        return [(r.dists[0, i], int(r.ids[0, i])) for i in range(int(k)) if r.ids[0, i] >= 0]
    L = max(int(k), int(beam_width))
    r = idx.search(as_f32(query_vector).reshape(1, -1), k=int(k), L=L, W=1, dist="pq" if pq else "exact", rerank=False,
                   sqrt_out=True, start=start)
    return [(r.dists[0, i], int(r.ids[0, i])) for i in range(int(k)) if r.ids[0, i] >= 0]


def beam_search(graph, query_vector, start_idx, beam_width=5, k=3, *, reference_semantics=False):
    return beam_search_with_pq(graph, query_vector, start_idx, beam_width, k, False, reference_semantics=reference_semantics)


_READER_CACHE = {}


def _reader_index(reader) -> GpuIndex:
    key = id(reader)
    hit = _READER_CACHE.get(key)
    if hit is not None and hit[0] is reader:
        return hit[1]
    if hasattr(reader, "records"):
        rec = reader.records()
    else:  # the reference's MMapNodeReader: same attributes, read the mapping directly
        rec = np.frombuffer(reader.mmap_obj, dtype=np.uint8)
    n = rec.size // (4 * (reader.D + reader.R))
    idx = GpuIndex.from_records(rec, n, reader.D, reader.R, medoid=0)
    if len(_READER_CACHE) > 8:
        old = _READER_CACHE.pop(next(iter(_READER_CACHE)))      # the oldest entry (dicts keep insertion order)
        old[1].close()
    _READER_CACHE[key] = (reader, idx)
    return idx


def beam_search_from_disk(reader, query_vector, start_id, beam_width=8, k=5):
    """Variant D (:719-760): exact search over the index.dat records with list size beam_width; returns the k
    best as (distance, id) with the non-squared distance like np.linalg.norm."""
    idx = _reader_index(reader)
    kk = min(int(k), int(beam_width))
    r = idx.search(as_f32(query_vector).reshape(1, -1), k=kk, L=int(beam_width), W=1, dist="exact", rerank=False, sqrt_out=True,
                   start=int(start_id))
    return [(r.dists[0, i], np.uint32(r.ids[0, i])) for i in range(kk) if r.ids[0, i] >= 0]


def search_batch(graph_or_reader, Q, k=10, L=100, W=1, use_pq=None, rerank=True, start_idx=None, **kw):
    """Additive batched entry point: Q f32[B,D] -> SearchResult(ids i32[B,k], dists f32[B,k], hops, visited, ...)."""
    if hasattr(graph_or_reader, "get_node"):
        idx = _reader_index(graph_or_reader)
        pq = False
    else:
        g = _as_graph(graph_or_reader)
        idx = g.gpu_index()
        pq = _pq_on(g) if use_pq is None else bool(use_pq)
        if start_idx is None:
            start_idx = g.medoid_idx or 0
    return idx.search(Q, k=k, L=L, W=W, dist="pq" if pq else "exact", rerank=rerank and pq,
                      start=None if start_idx is None else int(start_idx), **kw)
