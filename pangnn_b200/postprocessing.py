"""Output side of the path (``src/postprocessing.py:5-36``): ortholog groups from the
model's predictions = connected components of the edges predicted positive, computed on the device
(``ops.connected_components``), and the group table the reference calls ``holiest_of_all_tables.csv``.

The reference's ``write_groups_file`` is unused and does not merge sets as written (every pair is appended
again after the scan); the behaviour here is the intended one, one line per group of at least two genes."""
import os

import torch

from . import ops


def ortholog_groups(edge_index, binary_prediction, num_genes):
    """-> (labels [N] int32: smallest gene id of the gene's group, groups: list of gene-id lists with >= 2 genes,
    ordered by smallest id)."""
    labels = ops.connected_components(edge_index[0], edge_index[1], num_genes, select=binary_prediction)
    order = torch.argsort(labels.long(), stable=True)
    sl = labels[order]
    head = torch.ones_like(sl, dtype=torch.bool)
    head[1:] = sl[1:] != sl[:-1]
    starts = torch.nonzero(head).squeeze(1)
    sizes = torch.diff(torch.cat((starts, starts.new_tensor([sl.numel()]))))
    keep = sizes > 1
    order_h, starts_h, sizes_h = order.cpu().tolist(), starts[keep].cpu().tolist(), sizes[keep].cpu().tolist()
    return labels, [order_h[a:a + n] for a, n in zip(starts_h, sizes_h)]


def write_groups_file(groups, gene_ids_lst=None, path=os.path.join("data", "holiest_of_all_tables.csv")):
    """``group_<i>, gene, gene, ...`` per line (``src/postprocessing.py:31-36``; ids are positions when the
    dataset has no string ids, as for simulated data)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        for i, grp in enumerate(groups):
            names = [gene_ids_lst[g] if gene_ids_lst is not None else str(g) for g in grp]
            f.write(f"group_{i}, {', '.join(names)}\n")
    return path
