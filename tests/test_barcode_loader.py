"""C++ barcode-table loader (bdx_barcode_table_load) against the Python mirror of the
reference's preprocess_bc_file (fileio.jl:7-72) on the reference's own tables and on awkward inputs."""
import os

import pytest

import bdx_b200 as bdx
from bdx_b200 import capi


@pytest.mark.parametrize("name", ["demo1.tsv", "demo1.csv", "demo2.tsv", "demo2.csv"])
@pytest.mark.parametrize("complement,rev", [(False, False), (True, False), (False, True), (True, True)])
def test_reference_tables(refdata, name, complement, rev):
    path = os.path.join(refdata, "reference_files", name)
    assert capi.load_barcode_table(path, complement, rev) == bdx.preprocess_bc_file(path, complement, rev)


def _write(tmp_path, name, text):
    p = str(tmp_path / name)
    with open(p, "wb") as fh:
        fh.write(text)
    return p


def test_fasta(tmp_path):
    text = (b">bc1 first barcode\nACGTu\nacgn\n>bc2\tx\r\nTTTTNNAA\r\n>empty_record\n>bc3\n  GGCC  \n\nAA")
    for ext in ("bc.fasta", "bc.FA"):
        p = _write(tmp_path, ext, text)
        for c, r in ((False, False), (True, True)):
            got = capi.load_barcode_table(p, c, r)
            assert got == bdx.preprocess_bc_file(p, c, r)
    seqs, lens, ids = capi.load_barcode_table(p, False, False)
    assert seqs == ["ACGTTACGN", "TTTTNNAA", "GGCCAA"] and lens == [8, 6, 6]
    assert ids == ["bc1", "bc2", "empty_record", "bc3"]      # the reference's ID/sequence skew for empty records


def test_tables_with_quotes_annotation_and_order(tmp_path):
    csv = (b'ID,extra,Full_annotation,Full_seq\r\n'
           b'"s,1",x,BBBB--BB,acguNNtt\r\n'
           b'\r\n'
           b's2,"multi\nline",--BBBB,GGacgu\n'
           b's3,y,BBB,ANN')
    p = _write(tmp_path, "t.csv", csv)
    want = (["ACGTTT", "ACGT", "ANN"], [6, 4, 1], ["s,1", "s2", "s3"])
    assert capi.load_barcode_table(p, False, False) == want
    assert bdx.preprocess_bc_file(p, False, False) == want
    assert capi.load_barcode_table(p, True, True) == bdx.preprocess_bc_file(p, True, True)
    assert capi.load_barcode_table(p, True, True)[0] == ["AAACGT", "ACGT", "NNT"]
    tsv = b"Full_seq\tID\tFull_annotation\nACGT\tq\tBBBB\n"
    p = _write(tmp_path, "t.txt", tsv)       # anything that is not .csv is tab-separated (fileio.jl:35)
    assert capi.load_barcode_table(p, False, False) == (["ACGT"], [4], ["q"])


def test_errors(tmp_path):
    p = _write(tmp_path, "bad.csv", b"ID,Full_seq,Full_annotation\nz9,ACGT,BBB\n")
    with pytest.raises(capi.BdxError, match="Length mismatch between sequence and annotation for ID: z9"):
        capi.load_barcode_table(p, False, False)
    with pytest.raises(ValueError, match="Length mismatch between sequence and annotation for ID: z9"):
        bdx.preprocess_bc_file(p, False, False)
    p = _write(tmp_path, "nocol.tsv", b"ID\tSeq\nq\tACGT\n")
    with pytest.raises(capi.BdxError, match="needs columns"):
        capi.load_barcode_table(p, False, False)
    with pytest.raises(capi.BdxError, match="cannot open"):
        capi.load_barcode_table(str(tmp_path / "missing.tsv"), False, False)


def test_config_from_loader_matches_mirror(refdata):
    """The loader output feeds bdx_params exactly like the mirror's."""
    path = os.path.join(refdata, "reference_files", "demo2.csv")
    seqs, lens, ids = capi.load_barcode_table(path, True, True)
    cfg = bdx.DemuxConfig(bc_seqs=seqs, bc_lengths_no_N=lens, ids=ids)
    c = capi.Config(cfg)
    assert c.params.set1.n_barcodes == len(seqs)
    c.close()
