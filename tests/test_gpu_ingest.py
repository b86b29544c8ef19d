"""f2: the device tokenizers (csrc/ingest.cu) against the line parsers of hostdata.py, which mirror the reference's
loaders (pacybara_precluster.py:47-63, pacybara_cluster.py:302-331) and are themselves pinned to the reference by
tests/test_cli.py.  Bit-exact: same reads, same bytes, same id order, same genotypes, same uptags."""
import gzip
import io
import os

import numpy as np
import pytest

from pacybara_b200 import hostdata, pipeline, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from pacybara_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def host_view(ra, names):
    """ReadArrays + names -> plain Python values that do not depend on id numbering or strides."""
    h = lambda a: a.cpu().numpy() if hasattr(a, "data_ptr") else np.asarray(a)
    seq, qual, length = h(ra.seq), h(ra.qual), h(ra.length)
    gp, gm, gq = h(ra.geno_ptr), h(ra.geno_mut), h(ra.geno_q)
    R = length.shape[0]
    reads = [(seq[i, :length[i]].tobytes(), qual[i, :length[i]].tobytes()) for i in range(R)]
    assert not seq[np.arange(seq.shape[1])[None, :] >= length[:, None]].any()        # zero padded
    geno = [[(names["mut_names"][m], int(q)) for m, q in zip(gm[gp[i]:gp[i + 1]], gq[gp[i]:gp[i + 1]])] for i in range(R)]
    rank = h(ra.mut_rank)
    used = sorted(set(gm.tolist()))
    mut_order = [names["mut_names"][m] for m in sorted(used, key=lambda m: rank[m])]
    up = None
    if ra.up_id is not None:
        up = [names["up_names"][u] if u >= 0 else None for u in h(ra.up_id).tolist()]
    return dict(ids=list(names["ids"]), reads=reads, lexrank=h(ra.lexrank).tolist(), geno=geno, mut_order=mut_order, up=up)


def both(ctx, fastq: str, geno: str, uptag=None):
    from pacybara_b200 import ingest
    fb, gb = fastq.encode(), geno.encode()
    ub = fb if uptag is fastq else (uptag.encode() if uptag is not None else None)
    dev = host_view(*ingest.arrays_from_bytes(ctx, fb, gb, ub))
    wrap = lambda b: io.TextIOWrapper(io.BytesIO(b)) if b is not None else None
    ref = host_view(*pipeline.arrays_from_text(wrap(fb), wrap(gb), wrap(ub), ctx))
    return dev, ref


def assert_same(dev, ref):
    for k in ref:
        assert dev[k] == ref[k], k


FQ = ("@r10 pos=17 len=8\nACGTACGT\n+\nIIIIIIII\n"
      "@r9 pos=17\nACGTACGA\n+\nIIII#III\n"
      "@r100\nTTTT\n+\n@III\n"                       # quality starting with '@': the reference drops the read
      "@r11 x\nACGTACGT  \r\n+\r\nIIIIII!I\t\n"       # CRLF, trailing blanks
      "@q/7/ccs\nGGGGGGGGGGGG\n+\nIIIIIIIIIIII")      # no newline at the end
GE = ('"r10","A5C:40;T7G:33"\n'
      '"zzz","A1C:1"\n'                               # not a read: ignored
      'r9,=\n'                                        # no quotes
      '"r11","A5C:12"\n'
      '"r11","G9T:50;A5C:2"\r\n'                      # the last record of a read wins
      '"q/7/ccs","100_101insAAC:+7;A5C:0",extra\n'
      '"r100","T1A:5"\n')


def test_small_with_quirks(ctx):
    dev, ref = both(ctx, FQ, GE)
    assert_same(dev, ref)
    assert dev["ids"] == ["r10", "r9", "r11", "q/7/ccs"]
    assert dev["geno"][2] == [("G9T", 50), ("A5C", 2)]
    assert dev["lexrank"] == [1, 3, 2, 0]               # string order: q/7/ccs < r10 < r11 < r9


def test_uptags(ctx):
    up = ("@r10 a\nAAAA\n+\nIIII\n"
          "@r9\nCCCC \n+\nIIII\n"
          "@r10\nGGGG\n+\nIIII\n"                     # later record wins
          "@nobody\nTTTT\n+\nIIII\n"
          "@q/7/ccs\n@r11\nCCCC\n+\nIIII\n")           # a header right after a header: q/7/ccs gets no tag
    dev, ref = both(ctx, FQ, GE, up)
    assert_same(dev, ref)
    assert dev["up"] == ["GGGG", "CCCC", "CCCC", None]
    dev, ref = both(ctx, FQ, GE, FQ)                   # the barcode file as its own uptag file (non-combo libraries)
    assert_same(dev, ref)
    dev, ref = both(ctx, FQ, GE, "")                   # no records: uptags stay off
    assert_same(dev, ref)
    assert dev["up"] is None


@pytest.mark.parametrize("cfg,scale,style", [("cfg1", 1.0, "plain"), ("cfg3", 0.01, "padded"), ("cfg2", 0.05, "plain")])
def test_synthetic_files(ctx, tmp_path, cfg, scale, style):
    from pacybara_b200 import ingest
    lib = synth.make_config(cfg, seed=5, scale=scale, id_style=style)
    paths = lib.write_files(str(tmp_path))
    ra, names = ingest.arrays_from_files(ctx, paths["bc"], paths["geno"], paths["uptag"])
    assert hasattr(ra.seq, "data_ptr")                 # the tokenizers ran (no fallback)
    with hostdata.open_text(paths["bc"]) as fq, hostdata.open_text(paths["geno"]) as ge, hostdata.open_text(paths["uptag"]) as up:
        ref = pipeline.arrays_from_text(fq, ge, up, ctx)
    assert_same(host_view(ra, names), host_view(*ref))
    # and the clusters of the two loaders are the same table
    out = pipeline.cluster_reads(ctx, ra)
    a = pipeline.table_from_outputs(out, ra, names)
    b = pipeline.table_from_outputs(pipeline.cluster_reads(ctx, ref[0]), *ref)
    assert a == b and len(a) > 0
    # the writer: rows formatted on the device from the slices of the input texts (pcb_concat_tokens) == the host loop's file
    text = pipeline.table_text_from_outputs(ctx, out, ra, names)
    want = "virtualBarcode,upBarcode,reads,size,geno\n" + "".join(f"{v},{u},{r},{s},{g}\n" for v, u, r, s, g in a)
    assert text == want.encode("ascii")
    assert pipeline.table_text_from_outputs(ctx, out, ref[0], ref[1]) is None          # line-parser route: no slices, host loop


@pytest.mark.parametrize("what,fastq,geno", [
    ("non-ascii", FQ.replace("pos=17 len=8", "pos=17 lén"), GE),
    ("lone CR", FQ.replace("@r9 pos=17\n", "@r9 pos=17\r"), GE),
    ("text before the first header", "junk\n" + FQ, GE),
    ("quality shorter than sequence", FQ.replace("IIII#III", "IIII#II"), GE),
    ("header without id", FQ.replace("@r9 pos=17", "@ r9"), GE),
    ("duplicate read id", FQ.replace("@r9 pos=17", "@r10"), GE),
    ("a lone '@' line after the last record", FQ + "\n@\n", GE),
    ("genotype without quality", FQ, GE.replace("A5C:12", "A5C")),
    ("quality not a number", FQ, GE.replace("A5C:12", "A5C:1e3")),
    ("line without comma", FQ, GE + "oops\n"),
    ("mutation twice in a read", FQ, GE.replace("G9T:50;A5C:2", "A5C:50;A5C:2")),
])
def test_declines_what_it_does_not_model(ctx, what, fastq, geno):
    from pacybara_b200 import ingest
    with pytest.raises(ingest.NeedsLineParser):
        ingest.arrays_from_bytes(ctx, fastq.encode(), geno.encode())


def test_missing_genotype_is_the_references_keyerror(ctx):
    from pacybara_b200 import ingest
    with pytest.raises(KeyError):
        ingest.arrays_from_bytes(ctx, FQ.encode(), GE.replace('"r10",', '"r1O",').encode())
    wrap = lambda s: io.StringIO(s)
    with pytest.raises(KeyError):
        pipeline.arrays_from_text(wrap(FQ), wrap(GE.replace('"r10",', '"r1O",')), None, ctx)


def test_fallback_gives_the_line_parsers_answer(ctx, tmp_path):
    from pacybara_b200 import ingest
    fq, ge = tmp_path / "bc.fastq.gz", tmp_path / "geno.csv"
    with gzip.open(fq, "wt") as fh:
        fh.write(FQ)
    ge.write_text(GE.replace("G9T:50;A5C:2", "A5C:50;A5C:2"))           # dict semantics: A5C -> 2
    said = []
    ra, names = ingest.arrays_from_files(ctx, str(fq), str(ge), None, log=said.append)
    assert said and "line parser" in said[0]
    assert host_view(ra, names)["geno"][2] == [("A5C", 2)]


def test_empty_and_tiny(ctx):
    from pacybara_b200 import ingest
    ra, names = ingest.arrays_from_bytes(ctx, b"", b"")
    assert ra.n_reads == 0 and names["ids"] == []
    dev, ref = both(ctx, "@a\nA\n+\nI\n", "a,=\n")
    assert_same(dev, ref)
    assert dev["geno"] == [[]]


def test_rank_strings_long_ids(ctx):
    from pacybara_b200.api import Text
    rng = np.random.default_rng(3)
    ids = []
    for i in range(5000):
        n = int(rng.integers(1, 40))
        ids.append("m64012_190920_173625/" + "".join(rng.choice(list("0123456789/abc_"), n)))
    ids = sorted(set(ids), key=lambda s: rng.random())
    raw = "\n".join(ids).encode()
    off = np.cumsum([0] + [len(s) + 1 for s in ids[:-1]]).astype(np.int64)
    ln = np.array([len(s) for s in ids], dtype=np.int32)
    with Text(ctx, raw) as t:
        rank = t.rank_strings(off, ln, int(ln.max()))
    assert np.array_equal(rank, hostdata.lexrank_of(ids))


def test_quirky_files_against_the_reference_itself(ctx):
    """tests/golden/x_textquirks: CRLF records, trailing blanks, an '@' quality, no final newline, unquoted / CRLF /
    three-column genotype lines, ghost reads, reads recorded twice -- run through the UNMODIFIED reference
    (oracle/make_golden.py --add-textquirks).  The tokenizers must take these files as they are (no line parser) and the
    clusters must be the reference's table."""
    from pacybara_b200 import canon, ingest
    from tests import util
    case = util.load_golden("x_textquirks")
    ra, names = ingest.arrays_from_bytes(ctx, case["fastq"].encode(), case["geno"].encode(), case["uptag"].encode())
    assert hasattr(ra.seq, "data_ptr")
    reads = hostdata.parse_fastq_for_precluster(io.StringIO(case["fastq"]))
    assert names["ids"] == reads.ids and len(reads.ids) < case["fastq"].count("\n+")        # the '@' qualities dropped reads
    for full in (True, False):
        out = pipeline.cluster_reads(ctx, ra, full_table=full, **case["params"])
        table = pipeline.table_from_outputs(out, ra, names)
        assert canon.canon_table(table) == util.golden_table(case)
        text = pipeline.table_text_from_outputs(ctx, out, ra, names)                   # the device-formatted file, CRLF ids and all
        assert text == ("virtualBarcode,upBarcode,reads,size,geno\n" + "".join(f"{v},{u},{r},{s},{g}\n" for v, u, r, s, g in table)).encode()
    # and the bucketing text the reference printed, from the tokenizer's matrices
    fq = ingest.fastq_from_bytes(ctx, case["fastq"].encode())
    res = ctx.bucket(fq.seq, fq.qual, fq.length)
    res = {k: (v.cpu().numpy() if hasattr(v, "data_ptr") else v) for k, v in res.items()}
    buf = io.StringIO()
    hostdata.write_preclust_fastq(buf, fq.ids, res["bucket_ptr"], res["bucket_reads"], res["bucket_seq"], res["bucket_len"], res["bucket_qsum"])
    assert buf.getvalue() == case["preclust"]


def random_odd_texts(rng):
    """A small FASTQ + genoExtract pair with the formatting accidents the loaders tolerate (and a few they do not)."""
    n = int(rng.integers(1, 30))
    pool = ["".join(rng.choice(list("ACGT"), int(rng.integers(1, 12)))) for _ in range(int(rng.integers(1, 5)))]
    muts = [f"{int(rng.integers(1, 3000))}{a}>{b}" for a, b in zip(rng.choice(list("ACGT"), 6), rng.choice(list("ACGT"), 6))]
    fq, ge = [], []
    for i in range(n):
        rid = f"m{int(rng.integers(100))}/{i}/ccs"
        seq = pool[int(rng.integers(len(pool)))]
        qual = "".join(chr(int(c)) for c in rng.integers(33, 127, len(seq)))
        if rng.random() < 0.08:
            qual = "@" + qual[1:]
        eol = "\r\n" if rng.random() < 0.2 else "\n"
        pad = str(rng.choice(["", "", " ", "\t", "  "]))
        head = f"@{rid}" + str(rng.choice(["", " pos=17", " pos=NA len=3"] * 100 + ["\tx"]))     # a tab belongs to the id
        rec = f"{head}{pad}{eol}{seq}{pad}{eol}+{eol}{qual}{eol}"
        if rng.random() < 0.05:
            rec += eol
        if rng.random() < 0.02:
            rec = rec.replace("+", "", 1)
        fq.append(rec)
        if rng.random() < 0.004:
            continue                                          # a read without genotype record
        k = int(rng.integers(0, 4))
        g = ";".join(f"{m}:{int(rng.integers(0, 93))}" for m in rng.choice(muts, k, replace=False)) if k else "="
        if rng.random() < 0.1:
            ge.append(f'"{rid}","1A>C:7"\n')
        q = '"' if rng.random() < 0.7 else ""
        tail = ",x" if rng.random() < 0.1 else ""
        ge.append(f"{q}{rid}{q},{q}{g}{q}{tail}" + str(rng.choice(["", " ", "\t"])) + ("\r\n" if rng.random() < 0.2 else "\n"))
        if rng.random() < 0.05:
            ge.append(f"ghost{i},5A>T:40\n")
    fastq = "".join(fq)
    if rng.random() < 0.3:
        fastq = fastq.rstrip("\r\n")
    return fastq, "".join(ge)


def test_random_odd_texts_match_the_line_parser(ctx):
    """Whatever the tokenizers accept they must read exactly as the line parser does (which oracle/fuzz_against_reference.py
    pins to the reference); whatever they do not model they must decline, never guess."""
    from pacybara_b200 import ingest
    rng = np.random.default_rng(17)
    taken = declined = failed_alike = 0
    for _ in range(250):
        fastq, geno = random_odd_texts(rng)
        fb, gb = fastq.encode(), geno.encode()
        wrap = lambda b: io.TextIOWrapper(io.BytesIO(b))
        try:
            ref, ref_err = host_view(*pipeline.arrays_from_text(wrap(fb), wrap(gb), wrap(fb), ctx)), None
        except Exception as e:                                # noqa: BLE001
            ref, ref_err = None, type(e)
        try:
            dev, dev_err = host_view(*ingest.arrays_from_bytes(ctx, fb, gb, fb)), None
        except ingest.NeedsLineParser:
            declined += 1
            continue
        except Exception as e:                                # noqa: BLE001
            dev, dev_err = None, type(e)
        if ref_err is not None or dev_err is not None:
            assert ref_err is dev_err, (ref_err, dev_err, fastq, geno)
            failed_alike += 1
            continue
        for k in ref:
            assert dev[k] == ref[k], (k, fastq, geno)
        taken += 1
    assert taken > 100, (taken, declined, failed_alike)


def test_concat_tokens_against_the_host_restatement(ctx):
    """pcb_concat_tokens (csrc/format.cu): random slices, separators, empty tokens, the capacity retry."""
    from pacybara_b200 import tablefmt
    rng = np.random.default_rng(11)
    src = rng.integers(32, 127, size=5000, dtype=np.uint8)
    n_tok = 700
    tok_off = rng.integers(0, 4900, size=n_tok).astype(np.int64)
    tok_len = np.minimum(rng.integers(0, 40, size=n_tok), 5000 - tok_off).astype(np.int32)
    item_tok = rng.integers(0, n_tok, size=20000).astype(np.int32)
    item_sep = rng.choice(np.array([0, 44, 124, 10], dtype=np.uint8), size=20000)
    want = tablefmt.concat_host(src, tok_off, tok_len, item_tok, item_sep)
    assert ctx.concat_tokens(src, tok_off, tok_len, item_tok, item_sep) == want
    assert ctx.concat_tokens(src, tok_off, tok_len, item_tok, item_sep, size_hint=10) == want      # too small a buffer: one retry
    assert ctx.concat_tokens(src, tok_off, tok_len, item_tok[:0], item_sep[:0]) == b""
    from pacybara_b200 import _lib
    with pytest.raises(_lib.PcbError):
        ctx.concat_tokens(src, tok_off + 6000, tok_len, item_tok, item_sep)                        # a token outside the source
