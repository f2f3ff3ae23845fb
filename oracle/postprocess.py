"""TEST INFRASTRUCTURE — CPU statement of the output side: ortholog groups as connected
components of the edges predicted positive.  The reference's ``write_groups_file`` (``src/postprocessing.py:5-36``)
is unused and, as written, never merges two sets (it re-appends every pair); this is the intended behaviour and
there is no reference vector for it — parity unpinned for this row, anchored on scipy's connected_components."""
import numpy as np


def component_labels(src, dst, select, num_nodes):
    """label[i] = smallest node id of i's component over the selected edges."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    m = np.asarray(select).astype(bool) if select is not None else np.ones(len(src), dtype=bool)
    s, d = np.asarray(src)[m], np.asarray(dst)[m]
    g = coo_matrix((np.ones(s.size, dtype=np.int8), (s, d)), shape=(num_nodes, num_nodes))
    _, comp = connected_components(g, directed=False)
    first = np.full(comp.max() + 1 if comp.size else 0, num_nodes, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(num_nodes))
    return first[comp].astype(np.int32)


def groups(labels):
    """Components with at least two genes, as sorted id lists, ordered by their smallest id."""
    labels = np.asarray(labels)
    order = np.argsort(labels, kind="stable")
    out, start = [], 0
    sl = labels[order]
    for i in range(1, sl.size + 1):
        if i == sl.size or sl[i] != sl[start]:
            if i - start > 1:
                out.append(order[start:i].tolist())
            start = i
    return out
