#!/usr/bin/env python3
"""Mint the golden ``q_score_vs_logit.csv`` by running the UNMODIFIED reference writer
(``/root/reference/src/plot.py:451-504``, ``plot_sim_score_vs_logit``) under the import shim of
``oracle/ref_shim.py`` (matplotlib / seaborn are inert mocks; pandas is real).  Build-container only.

    python tests/golden/make_golden_csv.py

Inputs: the reference's own whole graph of the bundled two-genome data (config 1: edge order, Q-score weights,
labels, both max-candidate baselines as ``UnionGraphDataset`` built them), seeded pseudo-logits and the
reference's own max-logit-candidate baseline for them (``src/helper.py:494-576``).  Freezes inputs
(``tests/golden/q_score_vs_logit_c1.npz``) and the file the reference wrote (``..._c1.csv``)."""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    os.environ.setdefault("PYTHONHASHSEED", "0")
    import numpy as np
    import torch
    from oracle import ref_shim
    work = tempfile.mkdtemp(prefix="pangnn_golden_csv_")
    ref = ref_shim.load_reference([], work)
    import src.plot as plot                                    # the reference's writer, unmodified
    torch.manual_seed(0); np.random.seed(0)
    args = ref.args
    ds = ref.dataset.UnionGraphDataset(args.annotation, args.similarity, args.ribap_groups, split=(0.7, 0.15, 0.01),
                                       categorical_nodes=False, calculate_baseline=True)
    g = ds.test[0]
    E = g.edge_index.shape[1]
    gen = torch.Generator().manual_seed(1)
    logits = (torch.randn(E, generator=gen) * 3 + g.y * 2 - 1).float()
    logit_base = ref.helper.calculate_logit_baseline_labels(g, ds.sim_score_dict, logits, ds.gene_str_ids_lst,
                                                             ds.gene_id_position_dict)
    plot.plot_sim_score_vs_logit(g.y, g.edge_attr, logits, g.edge_index, ds.gene_str_ids_lst,
                                 (ds.base_labels, ds.base_labels_raw), logit_base)
    shutil.copy(os.path.join(work, "q_score_vs_logit.csv"), os.path.join(HERE, "q_score_vs_logit_c1.csv"))
    genes = np.asarray(ds.gene_str_ids_lst)
    np.savez_compressed(os.path.join(HERE, "q_score_vs_logit_c1.npz"), edge_index=g.edge_index.numpy(),
                        edge_attr=g.edge_attr.numpy(), y=g.y.numpy(), logits=logits.numpy(), genes=genes,
                        base_labels=np.asarray(ds.base_labels, dtype=np.int64),
                        base_labels_raw=np.asarray(ds.base_labels_raw, dtype=np.int64),
                        logit_baseline=np.asarray(logit_base, dtype=np.int64))
    print("wrote", E, "rows")


if __name__ == "__main__":
    main()
