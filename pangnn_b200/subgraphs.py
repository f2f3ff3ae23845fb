"""Training-mode sub-graph extraction on the device (SURVEY.md §8 a9, a10 and the sub-graph half of a11):
one sub-graph per ortholog group as ``src/dataset.py:222-322`` builds them with ``get_connected_nodes`` /
``get_neighbour_graph`` (``src/helper.py:327-417``) and ``build_edge_index`` (``src/preprocessing.py:73-118``).

The reference walks the groups one by one over dict-of-dicts keyed by gene-id strings (and needs a process pool
for it).  Here ALL groups advance together as relational operations over (sub-graph id, gene) keys on the
device — expand through the CSR of the similarity edges, sort, unique, sorted-table lookup:

  a10  n-hop closure: frontier (g, gene) -> out-neighbours (g, target), minus what g already holds, n times;
  a9   window: every closure gene proposes its +-n neighbours in global order; genes new to the sub-graph get
       local ids after the closure genes, in first-encounter order; both directions, duplicates removed;
       similarity edges with BOTH endpoints in the sub-graph (window genes included), by (src, dst);
  a11  union assembly ``[nb ; sim]`` with weights ``[1... ; w]`` (``src/dataset.py:287-303``).

Local numbering is canonical (closure genes ascending, then window genes in first-encounter order; the
reference's is CPython set order).  The result is ONE packed arena per attribute — exactly what
``ops.PackedGraphs`` / ``DeviceLoader`` collate from — and ``GraphList`` gives the list-of-graphs view the
reference's API exposes (``dataset.train`` / ``.val`` / ``.test``).  The large sorts run on our radix sort.
"""
import numpy as np
import torch

from . import ops
from .data import Data


def _sort_keys(keys):
    """Stable ascending sort of non-negative int64 keys -> (sorted, permutation) on our radix sort."""
    n = keys.numel()
    if n == 0:
        return keys, torch.zeros(0, dtype=torch.int64, device=keys.device)
    bits = max(int(keys.max().item()).bit_length(), 1)
    sk, perm = ops.sort_pairs_u64(keys, None, key_bits=bits)
    return sk, perm.long()


def _unique_sorted(keys):
    sk, _ = _sort_keys(keys)
    if sk.numel() == 0:
        return sk
    head = torch.ones_like(sk, dtype=torch.bool)
    head[1:] = sk[1:] != sk[:-1]
    return sk[head]


def _lookup(table, query):
    """Positions of ``query`` keys in the sorted ``table`` -> (found mask, position)."""
    if table.numel() == 0:
        return torch.zeros_like(query, dtype=torch.bool), torch.zeros_like(query)
    pos = torch.searchsorted(table, query).clamp_(max=table.numel() - 1)
    return table[pos] == query, pos


def _expand(rowptr, nodes):
    """Out-edges of ``nodes`` (with repetition): -> (index into ``nodes`` per edge, edge position)."""
    a = rowptr[nodes]
    deg = rowptr[nodes + 1] - a
    rep = torch.repeat_interleave(torch.arange(nodes.numel(), device=nodes.device), deg)
    start = torch.cumsum(deg, 0) - deg
    epos = a[rep] + (torch.arange(rep.numel(), device=nodes.device) - start[rep])
    return rep, epos, deg


def _ptr(counts):
    p = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=p[1:])
    return p


class SubGraphArena:
    """All sub-graphs of a dataset packed per attribute on the device (+ host copies of the offset tables)."""

    def __init__(self, attrs, node_ptr, class_balance, device):
        self.packed = ops.PackedGraphs.from_packed(attrs, node_ptr, device)
        self.num_graphs = self.packed.num_graphs
        self.class_balance = class_balance

    def graph(self, i):
        """``Data`` view of sub-graph ``i`` (slices of the arena; no copy)."""
        p = self.packed
        g = Data()
        for k in p.tensor_keys:
            a = p.attrs[k]
            lo, hi = int(a["ptr"][i]), int(a["ptr"][i + 1])
            g.__dict__[k] = a["packed"][:, lo:hi] if a["rows"] == 2 else a["packed"][lo:hi]
        return g


class GraphList:
    """List-of-graphs view over an arena: what ``dataset.train`` / ``.val`` / ``.test`` / ``.data_lst`` hold in
    training mode.  Indexing materialises a ``Data`` view; ``DeviceLoader`` collates straight from the arena."""

    def __init__(self, arena, ids):
        self.arena, self.ids = arena, np.asarray(ids, dtype=np.int32)

    def __len__(self):
        return int(self.ids.size)

    def __getitem__(self, i):
        if isinstance(i, (slice, list, np.ndarray)):
            return GraphList(self.arena, self.ids[i])
        return self.arena.graph(int(self.ids[i]))

    def __iter__(self):
        return (self.arena.graph(int(i)) for i in self.ids)


def extract(src, dst, w, y, num_genes, n, groups, union, gff_is_subset=False, chunks=1):
    """``src, dst`` int [E] sorted by (src, dst), ``w, y`` fp32 [E] (device); ``groups``: iterable of gene-id
    arrays (host).  -> ``SubGraphArena`` (sub-graphs in group order, groups of one gene / without similarity
    edges / (subset data) with fewer similarity edges than genes skipped as ``src/dataset.py:230,247,252``).
    ``chunks`` = the reference's ``--cpus``: its class balance is the MEAN over the worker chunks
    ``groups[i::cpus]`` of each chunk's neg / pos ratio (``src/dataset.py:128,141-142,319``), not the global ratio."""
    dev = src.device
    N = int(num_genes)
    n_groups_all = len(groups)
    orig = [i for i, g in enumerate(groups) if len(g) > 1]          # position in the reference's group list
    groups = [np.asarray(g, dtype=np.int64) for g in groups if len(g) > 1]
    if not groups:
        raise ValueError("no ortholog group with more than one gene")
    Gk = len(groups)
    sizes = torch.as_tensor(np.asarray([g.size for g in groups], dtype=np.int64), device=dev)
    gid0 = torch.repeat_interleave(torch.arange(Gk, device=dev), sizes)
    members = torch.as_tensor(np.concatenate(groups), device=dev)
    srcl, dstl = src.long(), dst.long()
    rowptr = _ptr(torch.bincount(srcl, minlength=N))

    # ---- a10: n-hop closure over out-edges, all groups at once
    conn = _unique_sorted(gid0 * N + members)
    frontier = conn
    for _ in range(n):
        gf, nf = frontier // N, frontier % N
        rep, epos, _ = _expand(rowptr, nf)
        cand = _unique_sorted(gf[rep] * N + dstl[epos])
        found, _ = _lookup(conn, cand)
        new = cand[~found]
        if new.numel() == 0:
            break
        conn, _ = _sort_keys(torch.cat((conn, new)))
        frontier = new
    gid_c, node_c = conn // N, conn % N
    cnt_c = torch.bincount(gid_c, minlength=Gk)
    start_c = _ptr(cnt_c)[:-1]
    local_c = torch.arange(conn.numel(), device=dev) - start_c[gid_c]

    # ---- a9: windows in global order; new genes numbered in first-encounter order
    delta = torch.cat((torch.arange(-n, 0, device=dev), torch.arange(1, n + 1, device=dev)))
    j = node_c[:, None] + delta[None, :]
    valid = (j >= 0) & (j < N)
    gw = gid_c[:, None].expand_as(j)[valid]
    lw = local_c[:, None].expand_as(j)[valid]
    keyw = gw * N + j[valid]                                        # in encounter order (row-major)
    found, pos = _lookup(conn, keyw)
    mk = keyw[~found]
    mk_sorted, perm = _sort_keys(mk)                                # stable: first encounter first within a key
    head = torch.ones_like(mk_sorted, dtype=torch.bool)
    head[1:] = mk_sorted[1:] != mk_sorted[:-1]
    win_keys = mk_sorted[head]                                      # distinct (g, gene), ascending
    by_seq = torch.argsort(perm[head])                              # ... in first-encounter order (groups by g)
    gid_ws = win_keys[by_seq] // N
    cnt_w = torch.bincount(gid_ws, minlength=Gk)
    start_w = _ptr(cnt_w)[:-1]
    local_w = torch.empty_like(win_keys)
    local_w[by_seq] = cnt_c[gid_ws] + torch.arange(win_keys.numel(), device=dev) - start_w[gid_ws]
    _, posw = _lookup(win_keys, keyw)
    lj = torch.where(found, local_c[pos], local_w[posw] if win_keys.numel() else torch.zeros_like(pos))
    cnt_n = cnt_c + cnt_w                                           # nodes per sub-graph
    M = max(int(cnt_n.max().item()), 1)
    if Gk * M * M >= 2 ** 62:
        raise ValueError("sub-graphs too large for 64-bit edge keys")
    nb_keys = _unique_sorted(torch.cat(((gw * M + lw) * M + lj, (gw * M + lj) * M + lw)))
    nb_g, nb_s, nb_t = nb_keys // (M * M), (nb_keys // M) % M, nb_keys % M

    # ---- similarity edges with both endpoints inside (window genes included)
    gid_o = torch.cat((gid_c, win_keys // N))
    node_o = torch.cat((node_c, win_keys % N))
    local_o = torch.cat((local_c, local_w))
    rep, epos, deg = _expand(rowptr, node_o)
    has_out = torch.zeros(Gk, dtype=torch.bool, device=dev)
    has_out[gid_o[deg > 0]] = True
    ge, se = gid_o[rep], local_o[rep]
    kt = ge * N + dstl[epos]
    f1, p1 = _lookup(conn, kt)
    f2, p2 = _lookup(win_keys, kt)
    lt = torch.where(f1, local_c[p1], local_w[p2] if win_keys.numel() else torch.zeros_like(p2))
    inside = f1 | f2
    ge, se, lt, epos = ge[inside], se[inside], lt[inside], epos[inside]
    sk, sperm = _sort_keys((ge * M + se) * M + lt)
    sim_g, sim_s, sim_t = sk // (M * M), (sk // M) % M, sk % M
    sim_w, sim_y = w[epos[sperm]].float(), y[epos[sperm]].float()
    cnt_e = torch.bincount(sim_g, minlength=Gk)

    # ---- group filters (src/dataset.py:247,252-253)
    few = cnt_e < sizes
    if not gff_is_subset and bool((few & has_out).any()):
        raise AssertionError("fewer similarity edges than genes in the origin family (src/dataset.py:253)")
    keep = has_out & ~few
    if not bool(keep.any()):
        raise ZeroDivisionError("no sub-graph with similarity edges (src/dataset.py:319)")
    newid = torch.cumsum(keep.long(), 0) - 1
    Gn = int(keep.sum().item())

    def kept(g, *cols):
        m = keep[g]
        return (newid[g[m]],) + tuple(c[m] for c in cols)

    # nodes: closure genes then window genes, by local id
    node_ptr = _ptr(cnt_n[keep])
    og, on, ol = kept(gid_o, node_o, local_o)
    node_id = torch.empty(int(node_ptr[-1].item()), dtype=torch.int64, device=dev)
    node_id[node_ptr[og] + ol] = on
    sg, ss, st_, sw, sy = kept(sim_g, sim_s, sim_t, sim_w, sim_y)
    ng, ns, nt = kept(nb_g, nb_s, nb_t)
    sim_cnt, nb_cnt = torch.bincount(sg, minlength=Gn), torch.bincount(ng, minlength=Gn)
    sim_ptr, nb_ptr = _ptr(sim_cnt), _ptr(nb_cnt)
    sim_ei, nb_ei = torch.stack((ss, st_)), torch.stack((ns, nt))
    # class balance: mean over the worker chunks that hold at least one group of neg / pos (src/dataset.py:141-142,319)
    chunks = max(int(chunks), 1)
    chunk_of_graph = (torch.as_tensor(np.asarray(orig, dtype=np.int64), device=dev) % chunks)[keep]
    ce = chunk_of_graph[sg]
    pos_c = torch.zeros(chunks, dtype=torch.float64, device=dev).index_add_(0, ce, sy.double())
    tot_c = torch.bincount(ce, minlength=chunks).double()
    ratios = []
    for c in range(min(chunks, n_groups_all)):                      # the non-empty chunks groups[c::chunks]
        if float(pos_c[c]) == 0:
            raise ZeroDivisionError("a worker chunk has no positive edge (src/dataset.py:319)")
        ratios.append((float(tot_c[c]) - float(pos_c[c])) / float(pos_c[c]))
    class_balance = sum(ratios) / len(ratios)
    host = lambda t: t.cpu().numpy()
    attrs = {"x": (torch.ones(node_id.numel(), 1, device=dev), host(node_ptr)),
             "edge_index": (sim_ei, host(sim_ptr))}
    if union:                                                       # a11: [nb ; sim], [1... ; w]
        u_ptr = nb_ptr + sim_ptr
        upos_nb = u_ptr[ng] + (torch.arange(ng.numel(), device=dev) - nb_ptr[ng])
        upos_sim = u_ptr[sg] + nb_cnt[sg] + (torch.arange(sg.numel(), device=dev) - sim_ptr[sg])
        total = int(u_ptr[-1].item())
        u_ei = torch.empty(2, total, dtype=torch.int64, device=dev)
        u_ei[:, upos_nb], u_ei[:, upos_sim] = nb_ei, sim_ei
        u_w = torch.ones(total, dtype=torch.float32, device=dev)
        u_w[upos_sim] = sw
        attrs["edge_attr"] = (u_w, host(u_ptr))
        attrs["y"] = (sy, host(sim_ptr))
        attrs["union_edge_index"] = (u_ei, host(u_ptr))
    else:
        attrs["edge_attr"] = (sw, host(sim_ptr))
        attrs["y"] = (sy, host(sim_ptr))
        attrs["neighbour_edge_index"] = (nb_ei, host(nb_ptr))
    attrs["node_id"] = (node_id, host(node_ptr))
    return SubGraphArena(attrs, host(node_ptr), class_balance, dev)
