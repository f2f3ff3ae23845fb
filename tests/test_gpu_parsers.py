"""§8f rank 2 on the device: GFF3 annotation and RIBAP group table parsers (``pangnn_gff_parse_lines``,
``pangnn_tsv_lookup_columns``) against what the UNMODIFIED reference loaders returned for the committed excerpts of the
bundled files (``tests/golden/make_parser_fixtures.py``), and against the pandas loaders of this package on synthetic
files with the awkward cases (comments, CRLF, missing fields, NA spellings, no start gene, odd ids)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "parsers")
GFFS = ("Cga_08-1274-3_RENAMED.gff", "Cga_12-4358_RENAMED.gff")


def _device_genes(files):
    from pangnn_b200 import preprocessing as pp
    genes, hashes = [], []
    for f in files:
        ids, h = pp.load_gff_device(f, device=DEV)
        genes += ids
        hashes.append(h)
    return genes, torch.cat(hashes)


def test_gff_and_ribap_fixtures_match_the_reference_loaders():
    from pangnn_b200 import ops, preprocessing as pp
    e = np.load(os.path.join(FIX, "expected.npz"))
    genes, hashes = _device_genes([os.path.join(FIX, n) for n in GFFS])
    assert genes == list(e["genes"])
    assert np.array_equal(hashes.cpu().numpy().view(np.uint64), ops.fnv1a64(genes))
    table = ops.gene_id_table_from_hashes(hashes, DEV)
    ref_table = ops.GeneIdTable(genes, DEV)
    assert torch.equal(table.hash, ref_table.hash) and torch.equal(table.pos, ref_table.pos)
    group_of, groups, is_subset = pp.load_ribap_groups_device(os.path.join(FIX, "ribap.csv"), list(e["genome_names"]),
                                                              table, len(genes), device=DEV)
    assert np.array_equal(group_of, e["group_of"]) and is_subset == bool(e["is_subset"])
    assert len(groups) == int(e["num_rows"])
    pos = {g: i for i, g in enumerate(genes)}
    _, host_groups, _ = pp.load_ribap_groups(os.path.join(FIX, "ribap.csv"), list(e["genome_names"]), pos)
    assert groups == host_groups


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_bytes(text.encode())
    return str(p)


@pytest.mark.parametrize("crlf", [False, True])
@pytest.mark.parametrize("with_start", [True, False])
def test_gff_edge_cases_against_pandas(tmp_path, crlf, with_start):
    from pangnn_b200 import preprocessing as pp
    rows = ["##gff-version 3", "# a comment line", ""]
    for i in range(40):
        attr = f"ID=ABC_{i:05d};Name=g{i};product=thing {i}"
        if i == 17 and with_start:
            attr += ";gene=hemB"
        if i == 5:
            attr = f"ID=abc_{i:05d};note=lower-case id is dropped"            # fails [A-Z]+_[0-9]+
        if i == 9:
            attr = f"ID=XY_{i:05d}"                                            # no ';' at all
        fields = ["contig1", "Prodigal", "CDS", str(100 * i + 1), str(100 * i + 90), ".", "+", "0", attr]
        if i == 11:
            fields[5] = ""                                                     # missing field -> dropna
        if i == 12:
            fields[1] = "NA"                                                   # pandas NA spelling -> dropna
        if i == 21:
            fields = fields[:6]                                                # short record -> dropna
        line = "\t".join(fields)
        if i == 30:
            line += "  # trailing comment is cut by pandas"
        rows.append(line)
    rows += ["##FASTA", ">contig1", "ACGTACGTACGT", "TTTTGGGGCCCC"]
    text = ("\r\n" if crlf else "\n").join(rows) + ("" if crlf else "\n")
    path = _write(tmp_path, "t.gff", text)
    ids, h = pp.load_gff_device(path, device=DEV)
    assert ids == pp.load_gff(path) and len(ids) == h.numel() and len(ids) > 30


def test_ribap_edge_cases_against_pandas(tmp_path):
    from pangnn_b200 import ops, preprocessing as pp
    genes = [f"AAA_{i:04d}" for i in range(50)] + [f"BBB_{i:04d}" for i in range(50)]
    pos = {g: i for i, g in enumerate(genes)}
    lines = ["# produced by a tool", "Cluster_ID\tAnnotation\tG1\tOther\tG2"]
    for r in range(30):
        a = f"AAA_{r:04d}" if r % 7 else ""                                   # missing cell
        b = f"BBB_{(r * 3) % 50:04d}" if r % 5 else "NA"
        if r == 13:
            a = "ZZZ_9999"                                                     # a gene that is not in the annotation
        lines.append(f"group{r}\tsomething # not a comment char here\t{a}\tCCC_{r:04d}\t{b}".replace(" # not a comment char here", ""))
    lines.insert(10, "")                                                       # blank line inside the table
    path = _write(tmp_path, "r.csv", "\n".join(lines))                         # no trailing newline
    # BBB column repeats ids (r*3 % 50 collides for r and r + 50/..): keep only rows that do not collide
    table = ops.GeneIdTable(genes, DEV)
    try:
        ref = pp.load_ribap_groups(path, ["G1", "G2"], pos)
    except AssertionError:
        with pytest.raises(AssertionError):
            pp.load_ribap_groups_device(path, ["G1", "G2"], table, len(genes), device=DEV)
        return
    got = pp.load_ribap_groups_device(path, ["G1", "G2"], table, len(genes), device=DEV)
    assert np.array_equal(got[0], ref[0]) and got[1] == ref[1] and got[2] == ref[2]
