#!/usr/bin/env python3
"""Mint the parser fixtures (§8f rank 2): excerpts of the reference's bundled input files (two GFF3 annotations, the
RIBAP group table) and what the UNMODIFIED reference loaders (``/root/reference/src/preprocessing.py:159-193,329-367``)
return for them, run under ``oracle/ref_shim.py``.  Build-container only.

    python tests/golden/make_parser_fixtures.py

The excerpts keep the file structure that matters to the parsers: the ``##`` header, annotation records around the
``hemB`` start gene (rotation), the ``##FASTA`` section (one-field records that ``dropna`` removes but that count for
the rotation index), and the RIBAP rows that name a gene of the excerpts."""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "parsers")
DATA = "/root/reference/data"
GFFS = ("Cga_08-1274-3_RENAMED.gff", "Cga_12-4358_RENAMED.gff")


def excerpt(lines):
    fasta = next(i for i, l in enumerate(lines) if l.startswith("##FASTA"))
    hem = next(i for i, l in enumerate(lines) if "hemB" in l)
    keep = sorted(set(range(0, 45)) | set(range(hem - 6, hem + 9)) | set(range(fasta - 12, fasta + 25)))
    return [lines[i] for i in keep if i < len(lines)]


def main():
    import numpy as np
    from oracle import ref_shim
    os.makedirs(OUT, exist_ok=True)
    for name in GFFS:
        with open(os.path.join(DATA, name)) as fh:
            lines = fh.readlines()
        with open(os.path.join(OUT, name), "w") as fh:
            fh.writelines(excerpt(lines))
    work = tempfile.mkdtemp(prefix="pangnn_parser_fix_")
    ref = ref_shim.load_reference([], work)
    genes, genome_names = [], []
    per_file = {}
    for name in GFFS:
        ids = list(ref.preprocessing.load_gff(os.path.join(OUT, name)).index)
        per_file[name] = ids
        genes += ids
        genome_names.append(name.rsplit(".", 1)[0].replace("_RENAMED", ""))
    gene_set = set(genes)
    with open(os.path.join(DATA, "holy_python_ribap_95.csv")) as fh:
        rows = fh.readlines()
    kept = [rows[0]] + [r for r in rows[1:] if gene_set & set(r.rstrip("\n").split("\t"))][:60] + rows[1:12]
    with open(os.path.join(OUT, "ribap.csv"), "w") as fh:
        fh.writelines(kept)
    # hit table: the rows of the bundled MMseqs2 table between two genes of the excerpts, plus rows with foreign ids
    with open(os.path.join(DATA, "mmseq2_result.csv")) as fh:
        hits = fh.readlines()
    both = [h for h in hits if h.split("\t", 2)[0] in gene_set and h.split("\t", 2)[1] in gene_set]
    foreign = [h for h in hits if h.split("\t", 2)[0] not in gene_set][:40]
    with open(os.path.join(OUT, "hits.tsv"), "w") as fh:
        fh.writelines(both[: len(both) // 2] + foreign + both[len(both) // 2:])
    rdict, rlst, is_subset = ref.preprocessing.load_ribap_groups(os.path.join(OUT, "ribap.csv"), genome_names)
    pos = {g: i for i, g in enumerate(genes)}
    group_of = np.full(len(genes), -1, dtype=np.int32)
    for gi, row in enumerate(rlst):
        for g in row:
            if g in pos:
                group_of[pos[g]] = gi
    np.savez_compressed(os.path.join(OUT, "expected.npz"), genes=np.asarray(genes),
                        counts=np.asarray([len(per_file[n]) for n in GFFS]), genome_names=np.asarray(genome_names),
                        group_of=group_of, num_rows=np.int64(len(rlst)), is_subset=np.bool_(is_subset))
    print({n: len(v) for n, v in per_file.items()}, "ribap rows", len(rlst), "labelled genes", int((group_of >= 0).sum()))


if __name__ == "__main__":
    main()
