"""CPU: rectorch_b200.data.DataReader / DatasetManager (CSV -> CSR ingest of libb200vae.so, host side) against
  (a) the known answers of the reference's own tests (tests/test_data.py:104-190 topn, :352-359 rated), and
  (b) fixtures produced by the unmodified reference DataReader (oracle/make_golden_data.py).
No GPU is needed: the ingest is host code; the matrices it returns are what DataSampler uploads.
"""
import json
import os

import numpy as np
import pytest
from scipy import sparse

from rectorch_b200._lib import B200VaeError
from rectorch_b200.configuration import DataConfig
from rectorch_b200.data import DataReader, DatasetManager, read_csv_csr

NAMES = ['train.csv', 'unique_iid.txt', 'unique_uid.txt', 'validation_tr.csv', 'validation_te.csv',
         'test_tr.csv', 'test_te.csv']


def _folder(tmp_path, files, cfg_extra):
    for n, f in zip(NAMES, files):
        with open(os.path.join(tmp_path, n), "w", newline="") as fh:
            fh.write(f)
    cfg = {"data_path": "NOT USED", "proc_path": str(tmp_path), "seed": 42, "threshold": 2.5, "separator": " ",
           "u_min": 1, "i_min": 1, "heldout": 1, "test_prop": 0.5}
    cfg.update(cfg_extra)
    p = os.path.join(tmp_path, "cfg.json")
    json.dump(cfg, open(p, "w"))
    return p


def test_datareader_reference_known_answers_topn(tmp_path):
    files = ['uid,iid\n0,0\n0,1\n1,2\n1,1\n', '2\n5\n3\n', '2\n4\n1\n3\n', 'uid,iid\n2,0\n', 'uid,iid\n2,1\n',
             'uid,iid\n3,0\n', 'uid,iid\n3,1\n']
    cfgp = _folder(tmp_path, files, {"topn": 1})
    with pytest.raises(TypeError):
        DataReader(1)
    reader = DataReader(cfgp)
    reader2 = DataReader(DataConfig(cfgp))
    assert reader.n_items == 3
    assert reader.cfg == reader2.cfg
    with pytest.raises(ValueError):
        reader.load_data("training")
    sp_data = reader.load_data("full")
    sp_train = reader.load_data("train")
    sp_vtr, sp_vte = reader.load_data("validation")
    sp_ttr, sp_tte = reader.load_data("test")
    assert sp_data.dtype == np.float64 and sparse.isspmatrix_csr(sp_train)
    assert np.all(sp_data.data == np.ones(8))
    r, c = sp_data.nonzero()
    assert np.all(r == np.array([0, 0, 1, 1, 2, 2, 3, 3]))
    assert np.all(c == np.array([0, 1, 1, 2, 0, 1, 0, 1]))
    assert np.all(sp_train.data == np.ones(4))
    r, c = sp_train.nonzero()
    assert np.all(r == np.array([0, 0, 1, 1])) and np.all(c == np.array([0, 1, 1, 2]))
    for m, col in ((sp_vtr, 0), (sp_vte, 1), (sp_ttr, 0), (sp_tte, 1)):
        assert np.all(m.data == np.array([1.]))
        r, c = m.nonzero()
        assert np.all(r == np.array([0])) and np.all(c == np.array([col]))
    # DatasetManager (data.py:522-560)
    man = DatasetManager(cfgp)
    assert man.n_items == 3 and man.training_set[1] is None
    tr, te = man.get_train_and_test()
    assert tr.shape == (4, 3) and te.shape == (4, 3)
    assert np.array_equal(tr.toarray(), np.array([[1, 1, 0], [0, 1, 1], [1, 1, 0], [1, 0, 0]], dtype=float))
    assert np.array_equal(te.toarray(), np.array([[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 1, 0]], dtype=float))


def test_datareader_reference_known_answers_rated(tmp_path):
    files = ['uid,iid,2\n0,0,3\n0,1,4\n1,2,4\n1,1,4\n', '2\n5\n3\n', '2\n4\n1\n3\n', 'uid,iid,2\n2,0,5\n',
             'uid,iid,2\n2,1,4\n', 'uid,iid,2\n3,0,5\n', 'uid,iid,2\n3,1,4\n']
    cfgp = _folder(tmp_path, files, {})          # no "topn" key: DefaultMunch -> None -> values from column 3
    sp_data = DataReader(cfgp).load_data("full")
    assert np.all(sp_data.data == np.array([3., 4., 4., 4., 5., 4., 5., 4.]))
    r, c = sp_data.nonzero()
    assert np.all(r == np.array([0, 0, 1, 1, 2, 2, 3, 3]))
    assert np.all(c == np.array([0, 1, 1, 2, 0, 1, 0, 1]))


def _same(m, g, key):
    m = m.tocsr()
    m.sum_duplicates()
    m.sort_indices()
    assert tuple(m.shape) == tuple(g[key + "/shape"]), key
    assert np.array_equal(m.indptr, g[key + "/indptr"]), key
    assert np.array_equal(m.indices, g[key + "/indices"]), key
    assert np.array_equal(m.data, g[key + "/data"]), key


@pytest.mark.parametrize("name", ["topn", "rated_dups", "crlf", "no_final_newline"])
def test_datareader_matches_reference_fixture(name, tmp_path, golden_dir):
    g = np.load(os.path.join(golden_dir, "datareader_%s.npz" % name))
    files = [str(g["file/" + n]) for n in NAMES]
    cfgp = _folder(tmp_path, files, {"topn": int(g["topn"])})
    reader = DataReader(cfgp)
    assert reader.n_items == int(g["n_items"])
    tr = reader.load_data("train")
    assert tr.has_canonical_format
    _same(tr, g, "train")
    a, b = reader.load_data("validation")
    _same(a, g, "validation_tr")
    _same(b, g, "validation_te")
    a, b = reader.load_data("test")
    _same(a, g, "test_tr")
    _same(b, g, "test_te")
    _same(reader.load_data("full"), g, "full")


def test_ingest_large_file_parallel_chunks(tmp_path):
    """A file large enough to be cut into several per-thread chunks parses to the same matrix as numpy."""
    rng = np.random.default_rng(0)
    n, n_users, n_items = 400_000, 20_000, 3_000
    u = rng.integers(0, n_users, n)
    i = rng.integers(0, n_items, n)
    v = rng.integers(1, 11, n) / 2.0
    p = os.path.join(tmp_path, "big.csv")
    with open(p, "w") as fh:
        fh.write("uid,iid,rating\n")
        fh.write("\n".join("%d,%d,%r" % t for t in zip(u.tolist(), i.tolist(), v.tolist())))
        fh.write("\n")
    assert os.path.getsize(p) > 4 << 20
    m = read_csv_csr(p, n_items, topn=False, n_rows=n_users)
    ref = sparse.csr_matrix((v, (u, i)), shape=(n_users, n_items), dtype=np.float64)
    ref.sum_duplicates()
    ref.sort_indices()
    assert np.array_equal(m.indptr, ref.indptr) and np.array_equal(m.indices, ref.indices)
    assert np.allclose(m.data, ref.data, rtol=0, atol=1e-12)
    ones = read_csv_csr(p, n_items, topn=True, n_rows=n_users)
    assert np.array_equal(ones.indptr, ref.indptr) and ones.data.sum() == n


def test_ingest_errors(tmp_path):
    p = os.path.join(tmp_path, "bad.csv")
    with open(p, "w") as fh:
        fh.write("uid,iid\n0,1\n1,x\n")
    with pytest.raises(B200VaeError, match="malformed record near data line 3"):
        read_csv_csr(p, 5)
    with pytest.raises(B200VaeError, match="cannot open"):
        read_csv_csr(os.path.join(tmp_path, "missing.csv"), 5)
    with open(p, "w") as fh:
        fh.write("uid,iid\n0,7\n")
    with pytest.raises(B200VaeError, match="column index 7"):
        read_csv_csr(p, 5)
    with open(p, "w") as fh:
        fh.write("uid,iid\n0,1\n")
    with pytest.raises(B200VaeError, match="no value column"):
        read_csv_csr(p, 5, topn=False)
    with open(p, "w") as fh:
        fh.write("uid,iid\n")
    m = read_csv_csr(p, 5)
    assert m.shape == (0, 5) and m.nnz == 0


def test_ingest_random_files_match_scipy(tmp_path):
    """Property check: random small files (duplicates, unsorted records, blank lines, '3.0'-style ids, spaces)
    parse to exactly what scipy builds from the same triples."""
    rng = np.random.default_rng(7)
    for case in range(25):
        n_users, n_items = int(rng.integers(1, 12)), int(rng.integers(1, 9))
        n = int(rng.integers(0, 40))
        u = rng.integers(0, n_users, n)
        i = rng.integers(0, n_items, n)
        v = rng.integers(-4, 9, n) / 4.0
        rated = bool(case % 2)
        lines = []
        for a, b, c in zip(u.tolist(), i.tolist(), v.tolist()):
            ua = ("%d.0" % a) if case % 5 == 0 else ("%d" % a)
            rec = "%s, %d" % (ua, b) if case % 7 == 0 else "%s,%d" % (ua, b)
            lines.append(rec + ((",%r" % c) if rated else ""))
            if case % 3 == 0 and rng.random() < 0.2:
                lines.append("")
        p = os.path.join(tmp_path, "f%d.csv" % case)
        with open(p, "w") as fh:
            fh.write("uid,iid,rating\n" if rated else "uid,iid\n")
            fh.write("\n".join(lines))
            if case % 4:
                fh.write("\n")
        m = read_csv_csr(p, n_items, topn=not rated, n_rows=n_users)
        vals = v if rated else np.ones(n)
        ref = sparse.csr_matrix((vals, (u, i)), shape=(n_users, n_items), dtype=np.float64)
        ref.sum_duplicates()
        ref.sort_indices()
        assert m.shape == ref.shape
        assert np.array_equal(m.indptr, ref.indptr) and np.array_equal(m.indices, ref.indices), case
        assert np.allclose(m.data, ref.data, rtol=0, atol=1e-12), case


RAW = ("1 1 4\n1 2 5\n1 3 2\n1 5 4\n" "2 2 3\n2 3 1\n2 5 4\n" "3 1 5\n3 2 5\n3 4 3\n3 5 4\n" "4 1 1\n4 3 4\n4 4 2\n4 5 4\n")


def _process(tmp_path, raw, cfg_extra, sep=" "):
    from rectorch_b200.data import DataProcessing
    rp = os.path.join(tmp_path, "raw.txt")
    with open(rp, "w") as fh:
        fh.write(raw)
    out = os.path.join(tmp_path, "proc")
    cfg = {"data_path": rp, "proc_path": out, "seed": 42, "threshold": 2.5, "separator": sep, "u_min": 1, "i_min": 1,
           "heldout": 1, "test_prop": 0.5}
    cfg.update(cfg_extra)
    cp = os.path.join(tmp_path, "cfg.json")
    json.dump(cfg, open(cp, "w"))
    dp = DataProcessing(cp)
    dp.process()
    return dp, cp, {n: open(os.path.join(out, n)).read() for n in os.listdir(out)}


def test_DataProcessing_reference_known_answers_topn(tmp_path):
    """Exact file contents expected by rectorch/tests/test_data.py:14-101 (they depend on the numpy RNG call
    sequence: user permutation and the per-user train/test choice)."""
    from rectorch_b200.data import DataProcessing
    with pytest.raises(TypeError):
        DataProcessing(1)
    dp, cp, files = _process(tmp_path, RAW, {"topn": 1})
    assert set(files) == {'validation_te.csv', 'validation_tr.csv', 'unique_iid.txt', 'unique_uid.txt', 'test_tr.csv',
                          'test_te.csv', 'train.csv'}
    assert files['train.csv'] == 'uid,iid\n0,0\n0,1\n1,2\n1,1\n'
    assert files['unique_iid.txt'] == '2\n5\n3\n'
    assert files['unique_uid.txt'] == '2\n4\n1\n3\n'
    assert files['validation_tr.csv'] == 'uid,iid\n2,0\n' and files['validation_te.csv'] == 'uid,iid\n2,1\n'
    assert files['test_tr.csv'] == 'uid,iid\n3,0\n' and files['test_te.csv'] == 'uid,iid\n3,1\n'
    assert DataProcessing(DataConfig(cp)).cfg == dp.cfg
    assert dp.u2id == {2: 0, 4: 1, 1: 2, 3: 3} and dp.i2id == {2: 0, 5: 1, 3: 2}
    sp = DataReader(cp).load_data("full")             # and the reader takes it from there (test_data.py:153-162)
    r, c = sp.nonzero()
    assert np.all(r == np.array([0, 0, 1, 1, 2, 2, 3, 3])) and np.all(c == np.array([0, 1, 1, 2, 0, 1, 0, 1]))


def test_DataProcessing_reference_known_answers_rated(tmp_path):
    """rectorch/tests/test_data.py:280-359 (no ``topn`` key: the rating column is kept, named after its position)."""
    _, cp, files = _process(tmp_path, RAW, {})
    assert files['train.csv'] == 'uid,iid,2\n0,0,3\n0,1,4\n1,2,4\n1,1,4\n'
    assert files['validation_tr.csv'] == 'uid,iid,2\n2,0,5\n' and files['validation_te.csv'] == 'uid,iid,2\n2,1,4\n'
    assert files['test_tr.csv'] == 'uid,iid,2\n3,0,5\n' and files['test_te.csv'] == 'uid,iid,2\n3,1,4\n'
    sp = DataReader(cp).load_data("full")
    assert np.all(sp.data == np.array([3., 4., 4., 4., 5., 4., 5., 4.]))


def test_DataProcessing_pipeline_invariants(tmp_path):
    """A larger file with a header, string user ids and float ratings: every step's contract holds and the reader
    can load what was written."""
    rng = np.random.default_rng(3)
    n_users, n_items = 120, 60
    lines = ["user,item,rating,ts"]
    for u in range(n_users):
        for i in rng.choice(n_items, size=int(rng.integers(1, 15)), replace=False):
            lines.append("u%03d,%d,%s,%d" % (u, 100 + i, repr(float(rng.integers(1, 11)) / 2), 1000 + u))
    dp, cp, files = _process(tmp_path, "\n".join(lines) + "\n",
                             {"header": 0, "threshold": 1.0, "u_min": 3, "i_min": 2, "heldout": 15, "test_prop": 0.2,
                              "seed": 7}, sep=",")
    uids = files['unique_uid.txt'].split()
    iids = files['unique_iid.txt'].split()
    assert len(set(uids)) == len(uids) and all(u.startswith("u") for u in uids) and len(set(iids)) == len(iids)
    assert files['train.csv'].splitlines()[0] == "uid,iid,rating,ts"
    reader = DataReader(cp)
    assert reader.n_items == len(iids)
    tr = reader.load_data("train")
    vtr, vte = reader.load_data("validation")
    ttr, tte = reader.load_data("test")
    n_tr = len(uids) - vtr.shape[0] - ttr.shape[0]
    assert tr.shape == (n_tr, len(iids)) and n_tr == len(uids) - 2 * 15 + (15 - vtr.shape[0]) + (15 - ttr.shape[0])
    assert np.diff(tr.indptr).min() >= 3                                   # u_min survived the item filter order
    for a, b in ((vtr, vte), (ttr, tte)):
        assert a.shape == b.shape and a.shape[0] <= 15
        la, lb = np.diff(a.indptr), np.diff(b.indptr)
        assert la.min() >= 1 and lb.min() >= 1                            # at least one item on each side
        assert np.all(lb == np.maximum((0.2 * (la + lb)).astype(int), 1))  # test_prop with the at-least-one rule
        assert a.multiply(b).nnz == 0                                       # the two parts are disjoint
    assert tr.data.min() > 1.0                                              # threshold applied (values kept: no topn)
    full = reader.load_data("full")
    assert full.shape == (len(uids), len(iids)) and full.nnz == tr.nnz + vtr.nnz + vte.nnz + ttr.nnz + tte.nnz


def test_DataProcessing_pandas_and_plain_paths_write_the_same_files(tmp_path, monkeypatch):
    """Reading / writing go through pandas' C parser and writer when pandas is importable and through plain Python
    otherwise: both must produce byte-identical output (integer ids, float ratings, a string column, a header)."""
    import sys
    rng = np.random.default_rng(5)
    lines = ["user,item,rating,tag"]
    for u in range(80):
        for i in rng.choice(50, size=int(rng.integers(2, 12)), replace=False):
            lines.append("%d,%d,%s, t%d" % (1000 + u, i, repr(float(rng.integers(1, 11)) / 2), u % 3))
    raw = "\n".join(lines) + "\n"
    extra = {"header": 0, "threshold": 1.0, "u_min": 2, "i_min": 2, "heldout": 10, "test_prop": 0.25, "seed": 3}
    os.makedirs(os.path.join(tmp_path, "a"))
    os.makedirs(os.path.join(tmp_path, "b"))
    _, _, with_pandas = _process(os.path.join(tmp_path, "a"), raw, extra, sep=",")
    monkeypatch.setitem(sys.modules, "pandas", None)            # `import pandas` now raises ImportError
    _, _, plain = _process(os.path.join(tmp_path, "b"), raw, extra, sep=",")
    assert set(with_pandas) == set(plain)
    for name in plain:
        assert with_pandas[name] == plain[name], name
