"""Generate tests/golden/datareader_*.npz by running the UNMODIFIED reference DataReader
(/root/reference/rectorch/data.py:328-420: pandas.read_csv + scipy csr_matrix) on synthetic
pre-processed folders.  Run in the build container:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_data.py

Each fixture holds the text of the seven files of a ``proc_path`` folder and the matrices the reference
returns for 'train', 'validation', 'test' and 'full' (indptr / indices / data / shape of the canonical CSR).
tests/test_data.py writes the files back to a temporary folder and requires rectorch_b200.data.DataReader to
return the same matrices.  The cases cover: topn (all-ones) and rated files, duplicate (uid, iid) records
(summed), unsorted records, users that only appear in the *_te file (dropped), CRLF line ends, a trailing
blank line and a file without a final newline.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from rectorch.data import DataReader          # noqa: E402  (reference)

OUT = os.path.join(ROOT, "tests", "golden")
NAMES = ['train.csv', 'unique_iid.txt', 'unique_uid.txt', 'validation_tr.csv', 'validation_te.csv',
         'test_tr.csv', 'test_te.csv']


def csv_text(rng, users, n_items, rated, nl="\n", dup_rate=0.0, final_nl=True, blank_tail=False, mean_len=6):
    lines = []
    for u in users:
        k = int(max(1, rng.poisson(mean_len)))
        items = rng.choice(n_items, size=min(k, n_items), replace=False)
        recs = [(u, int(i)) for i in items]
        for (uu, ii) in list(recs):
            if rng.random() < dup_rate:
                recs.append((uu, ii))
        for (uu, ii) in recs:
            if rated:
                lines.append("%d,%d,%s" % (uu, ii, repr(float(rng.integers(1, 11)) / 2.0)))
            else:
                lines.append("%d,%d" % (uu, ii))
    order = rng.permutation(len(lines))
    lines = [lines[i] for i in order]
    head = "uid,iid,rating" if rated else "uid,iid"
    text = nl.join([head] + lines)
    if final_nl:
        text += nl
    if blank_tail:
        text += nl
    return text


def make_case(name, seed, n_train, n_val, n_test, n_items, rated, **kw):
    rng = np.random.default_rng(seed)
    tr_users = list(range(n_train))
    va_users = list(range(n_train, n_train + n_val))
    te_users = list(range(n_train + n_val, n_train + n_val + n_test))
    texts = {
        'train.csv': csv_text(rng, tr_users, n_items, rated, **kw),
        'unique_iid.txt': "".join("%d\n" % (1000 + i) for i in range(n_items)),
        'unique_uid.txt': "".join("%d\n" % (5000 + u) for u in range(n_train + n_val + n_test)),
        # one validation user has no training part at all (must be dropped), one has no test part
        'validation_tr.csv': csv_text(rng, va_users[:-1], n_items, rated, **kw),
        'validation_te.csv': csv_text(rng, va_users[1:], n_items, rated, mean_len=2, **{k: v for k, v in kw.items() if k != "mean_len"}),
        'test_tr.csv': csv_text(rng, te_users, n_items, rated, **kw),
        'test_te.csv': csv_text(rng, te_users, n_items, rated, mean_len=2, **{k: v for k, v in kw.items() if k != "mean_len"}),
    }
    out = {"topn": np.int64(0 if rated else 1), "n_items": np.int64(n_items)}
    with tempfile.TemporaryDirectory() as d:
        for n in NAMES:
            with open(os.path.join(d, n), "w", newline="") as fh:
                fh.write(texts[n])
            out["file/" + n] = np.array(texts[n])
        cfg = {"proc_path": d, "topn": 0 if rated else 1}
        cfgp = os.path.join(d, "cfg.json")
        json.dump(cfg, open(cfgp, "w"))
        reader = DataReader(cfgp)
        assert reader.n_items == n_items

        def put(key, m):
            m = m.tocsr()
            m.sum_duplicates()
            m.sort_indices()
            out[key + "/indptr"] = m.indptr.astype(np.int64)
            out[key + "/indices"] = m.indices.astype(np.int32)
            out[key + "/data"] = m.data.astype(np.float64)
            out[key + "/shape"] = np.array(m.shape, dtype=np.int64)

        put("train", reader.load_data("train"))
        a, b = reader.load_data("validation")
        put("validation_tr", a)
        put("validation_te", b)
        a, b = reader.load_data("test")
        put("test_tr", a)
        put("test_te", b)
        put("full", reader.load_data("full"))
    np.savez_compressed(os.path.join(OUT, "datareader_%s.npz" % name), **out)
    print(name, {k: v.tolist() for k, v in out.items() if k.endswith("/shape")})


if __name__ == "__main__":
    make_case("topn", 1, 40, 6, 5, 37, rated=False)
    make_case("rated_dups", 2, 30, 5, 5, 23, rated=True, dup_rate=0.3)
    make_case("crlf", 3, 12, 4, 4, 11, rated=True, nl="\r\n", dup_rate=0.1, blank_tail=True)
    make_case("no_final_newline", 4, 12, 4, 4, 11, rated=False, final_nl=False)
