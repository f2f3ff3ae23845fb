"""Input side of the hot path: file parsers -> device hit table -> normalised edge list.

Mirrors the reference's ``src/preprocessing.py`` API names where they exist, but carries arrays
(hit tables in node ids) instead of dict-of-dicts keyed by gene-id strings:

  load_gff                 src/preprocessing.py:329-367   (host, pandas; load_gff_device: parsed on the device)
  load_similarity_score    src/preprocessing.py:388-426   (host parse, scores min-centred)
  load_ribap_groups        src/preprocessing.py:159-193   (host parse -> group_of[node]; load_ribap_groups_device)
  normalize_sim_scores     src/preprocessing.py:370-385,430-548 + build_edge_index / map_edge_weights /
                           map_labels_to_edge_index       (DEVICE: sort, dedupe, trivial filter,
                           segmented softmax + Q-score, compaction — pangnn_hits_* in the C ABI)
  neighbour_band           src/dataset.py:351-366         (device index arithmetic)
"""
import os
import re

import numpy as np
import torch

from . import ops
from .setup import args, log

_GENE_ID = re.compile(r"[A-Z]+_[0-9]+")


def load_gff(annotation_file_name, start_gene="hemB"):
    """Gene ids of one genome in the reference's order: rotated to start at the first record whose
    attribute mentions ``start_gene``, incomplete records dropped, ids = ``ID=...`` up to ';',
    kept only when they look like ``[A-Z]+_[0-9]+``."""
    import pandas as pd
    cols = ["seqname", "source", "feature", "start", "end", "score", "strand", "frame", "attribute"]
    df = pd.read_csv(annotation_file_name, comment="#", sep="\t", names=cols,
                     dtype={c: str for c in cols if c not in ("start", "end")} | {"start": "Int64", "end": "Int64"})
    hits = df.index[df["attribute"].str.contains(start_gene, na=False)].tolist()
    if hits:
        start = hits[0]
    else:
        log.error(f"Could not find start gene '{start_gene}' in annotation file {annotation_file_name}.")
        start = 1
    df = pd.concat([df.iloc[start:, :], df.iloc[:start, :]]).reset_index(drop=True).dropna()
    ids = df["attribute"].str.replace(";.*", "", regex=True).str.replace("ID=", "", regex=True)
    return [g for g in ids.tolist() if _GENE_ID.search(g)]


def genome_name_of(gff_file):
    return os.path.basename(gff_file).rsplit(".", 1)[0].replace("_RENAMED", "")   # src/dataset.py:96


def load_similarity_score(similarity_score_file, gene_id_position_dict, center_scores=True):
    """MMseqs2 16-column TSV -> (q, t, bits) int32/int32/float64 in FILE ORDER, restricted to rows
    whose query and target are both known genes, scores shifted to ``bits - min + 1``."""
    import pandas as pd
    names = ["query", "target", "pident", "alnlen", "mismatch", "gapopen", "qstart", "qend", "qlen",
             "tstart", "tend", "tlen", "qcov", "tcov", "evalue", "bits"]
    df = pd.read_csv(similarity_score_file, comment="#", sep="\t", names=names, usecols=["query", "target", "bits"])
    q = df["query"].map(gene_id_position_dict)
    t = df["target"].map(gene_id_position_dict)
    keep = q.notna() & t.notna()
    q, t = q[keep].to_numpy(np.int64), t[keep].to_numpy(np.int64)
    bits = df["bits"][keep].to_numpy(np.float64)
    if center_scores and bits.size:
        bits = bits - bits.min() + 1
    return q.astype(np.int32), t.astype(np.int32), bits


def load_similarity_score_device(similarity_score_file, gene_str_ids_lst, center_scores=True, device="cuda", table=None):
    """``load_similarity_score`` with the parse on the device (``ops.parse_hits_tsv``): the file's bytes are
    copied to HBM once; ids are matched by hash against the table of known genes; same result
    ((q, t, bits) in file order, rows with an unknown id dropped, ``bits - min + 1``) as device tensors."""
    raw = np.fromfile(similarity_score_file, dtype=np.uint8)
    text = torch.from_numpy(raw).to(device)
    q, t, bits = ops.parse_hits_tsv(text, table if table is not None else ops.GeneIdTable(gene_str_ids_lst, device))
    keep = (q >= 0) & (t >= 0)
    q, t, bits = q[keep], t[keep], bits[keep]
    if center_scores and bits.numel():
        bits = bits - bits.min() + 1
    return q, t, bits


def load_gff_device(annotation_file_name, start_gene="hemB", device="cuda"):
    """``load_gff`` with the parse on the device (``ops.parse_gff``): -> (gene ids as strings in the reference's
    order, their FNV-1a 64 hashes as a device tensor).  The strings are sliced out of the file's bytes in one
    vectorised gather; the hashes feed ``ops.gene_id_table_from_hashes`` without touching the strings again."""
    raw = np.fromfile(annotation_file_name, dtype=np.uint8)
    h, off, ln = ops.parse_gff(torch.from_numpy(raw).to(device), start_gene)
    return ops.gather_strings(raw, off.cpu().numpy(), ln.cpu().numpy()), h


def load_ribap_groups_device(ribap_group_file, genome_name_lst, table, num_genes, device="cuda"):
    """``load_ribap_groups`` with the table parsed on the device: every cell of the loaded genomes' columns is hashed
    and looked up in ``table`` (``ops.GeneIdTable``) by ``pangnn_tsv_lookup_columns``.
    -> (group_of [N] int32 numpy, groups as node-id arrays, is_subset) — same values as the host loader."""
    raw = np.fromfile(ribap_group_file, dtype=np.uint8)
    text = raw.tobytes()
    # header = first record line (pandas header=0 after comment removal); parsed here, it is one line
    pos, header = 0, None
    while pos < len(text):
        end = text.find(b"\n", pos)
        end = len(text) if end < 0 else end
        line = text[pos:end].split(b"#", 1)[0].rstrip(b"\r")
        pos = end + 1
        if line:
            header = [c.decode() for c in line.split(b"\t")]
            break
    if header is None:
        return np.full(num_genes, -1, dtype=np.int32), [], False
    wanted = set(genome_name_lst)
    col_slot, K = [], 0
    for c in header:
        col_slot.append(K if c in wanted else -1)
        K += c in wanted
    is_subset = any(c not in wanted for c in header)
    if K == 0:
        return np.full(num_genes, -1, dtype=np.int32), [], is_subset
    out, flag = ops.lookup_columns(torch.from_numpy(raw).to(device), table, col_slot, K)
    rec = torch.nonzero(flag).squeeze(1)[1:]                    # data records: every record line after the header
    mat = out[rec]
    rows = torch.arange(mat.size(0), device=mat.device, dtype=torch.int32).unsqueeze(1).expand_as(mat)
    known = mat >= 0
    ids, grp = mat[known].long(), rows[known]
    if ids.numel() and int(torch.bincount(ids, minlength=num_genes).max().item()) > 1:
        raise AssertionError("a gene belongs to more than one gene family (src/preprocessing.py:183)")
    group_of = torch.full((num_genes,), -1, dtype=torch.int32, device=mat.device)
    group_of[ids] = grp
    mat_h, known_h = mat.cpu().numpy(), known.cpu().numpy()
    groups = np.split(mat_h[known_h], np.cumsum(known_h.sum(1))[:-1]) if mat_h.shape[0] else []
    return group_of.cpu().numpy(), [g.tolist() for g in groups], is_subset


def load_ribap_groups(ribap_group_file, genome_name_lst, gene_id_position_dict):
    """RIBAP table -> (group_of [N] int32 with -1 = no group, groups as lists of node ids, is_subset).
    Only the columns of the loaded genomes are used; a gene may appear in one row only."""
    import pandas as pd
    df = pd.read_csv(ribap_group_file, comment="#", sep="\t", header=0)
    extra = df.columns.difference(genome_name_lst)
    is_subset = not extra.empty
    df = df.drop(columns=extra)
    N = len(gene_id_position_dict)
    group_of = np.full(N, -1, dtype=np.int32)
    groups = []
    for gi, row in enumerate(df.itertuples(index=False)):
        members = [gene_id_position_dict[g] for g in row if isinstance(g, str) and g in gene_id_position_dict]
        named = [g for g in row if isinstance(g, str)]
        for g in named:
            if g in gene_id_position_dict:
                pos = gene_id_position_dict[g]
                assert group_of[pos] == -1, f"{g} already belongs to a gene family"
                group_of[pos] = gi
        groups.append(members)
    return group_of, groups, is_subset


def normalize_sim_scores(q, t, bits, genome_of, group_of=None, num_nodes=None, t_norm=None,
                         epsilon=1e-8, pseudo_count=1.0, include_trivial=None, device="cuda"):
    """Hit table (host numpy or device tensors, any order, duplicates allowed) -> device tensors
    ``(src int32, dst int32, w fp32, y fp32)`` sorted by (src, dst).  One H2D copy, then everything
    (radix sort, dedupe-last, trivial-case filter, segmented softmax + Q-score, label lookup,
    compaction) runs in the CUDA library."""
    temp = args.normalization_temp if t_norm is None else t_norm
    if temp == 0:
        raise ValueError("normalization_temp == 0 (no normalisation) is not supported: the reference "
                         "never assigns sim_score_dict in that branch (src/dataset.py:111-114)")
    if include_trivial is None:
        include_trivial = args.include_trivial
    as_dev = lambda a, dt: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(
        device=device, dtype=dt, non_blocking=True)
    genome_of_d = as_dev(genome_of, torch.int32)
    N = int(num_nodes if num_nodes is not None else genome_of_d.numel())
    qs, ts, bs = ops.hits_sort_unique(as_dev(q, torch.int32), as_dev(t, torch.int32),
                                      as_dev(bits, torch.float64), N)
    group_d = as_dev(group_of, torch.int32) if group_of is not None else None
    return ops.hits_normalize(qs, ts, bs, genome_of_d, group_d, temp=temp, eps=epsilon,
                              pseudo=pseudo_count, drop_trivial=not include_trivial)


def neighbour_band(num_genes, n, device="cuda"):
    """Whole-graph neighbour edges i -> j, j in [i-n, i+n] ∩ [0, N), self loop included, crossing
    genome seams, in the reference's loop order (``src/dataset.py:356-361``): one closed-form kernel."""
    return ops.neighbour_band(num_genes, n, device)


def baseline_labels(src, dst, score, genome_of):
    """Max-candidate baseline (``src/helper.py:437-485``; with logits as ``score``: ``:494-576``): 1 iff
    no candidate of the same (query, target genome) segment scores strictly higher.  The table must be
    sorted by (src, dst) — every table this package produces is.  One segmented arg-max kernel."""
    return ops.segment_max_labels(src, dst, score, genome_of).to(torch.int64)
