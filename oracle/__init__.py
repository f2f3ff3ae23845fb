"""CPU oracle for the panGNN message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pangnn_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs do, and only as the checker / the CPU arm, never as the product path.

What it is: a plain numpy / torch-CPU restatement of the reference algorithm on the path named
by SURVEY.md section 8 (rows a1-a20), each function citing the reference ``file:line`` it follows.

How it is pinned:
  * rows a1-a11, a13, a16-a19 (preprocessing, graph assembly, embedding, edge scorer, loss) are
    pinned against the reference's OWN code: ``oracle/ref_shim.py`` imports ``/root/reference/src``
    unmodified (under import stubs for the absent ``torch_geometric`` / ``matplotlib`` /
    ``seaborn``), ``tests/golden/make_golden.py`` runs it and freezes inputs+outputs into
    ``tests/golden/*.npz``; ``tests/test_oracle_*.py`` compare this restatement to them.
  * row a14 (``torch_geometric.nn.GCNConv``) is a THIRD-PARTY dependency that is absent from
    ``/root/reference`` and un-pinned there (the ``pangnn.yaml`` the README points to is not in
    the repo).  ``oracle/gcn.py`` restates PyG's published ``gcn_norm`` + ``GCNConv.forward``
    (add_self_loops=False, normalize=True, bias=True, aggr='add', flow='source_to_target').
    The reference holds no test or golden vector for it: **parity unpinned** for that row —
    it is anchored on the reference call sites (``src/gnn.py:100-102,129-165``) and on
    hand-computed known-answer cases in ``tests/test_oracle_model.py``.
"""
