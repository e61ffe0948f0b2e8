"""Host ingest throughput: b200vae_csv_* (csrc/ingest.cu, multi-threaded mmap parse + counting sort) against the
reference's pd.read_csv + scipy.sparse.csr_matrix((values, (rows, cols))) (rectorch/data.py:375-391) on the same
train.csv-shaped file.  CPU only.

    python scripts/ingest_bench.py [--records 5000000]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200.data import read_csv_csr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=5_000_000)
ap.add_argument("--users", type=int, default=120_000)
ap.add_argument("--items", type=int, default=20_000)
args = ap.parse_args()

rng = np.random.default_rng(0)
u = np.sort(rng.integers(0, args.users, args.records))
i = rng.integers(0, args.items, args.records)
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "train.csv")
    t0 = time.perf_counter()
    with open(p, "w") as fh:
        fh.write("uid,iid\n")
        np.savetxt(fh, np.stack([u, i], 1), fmt="%d", delimiter=",")
    print("file: %d records, %.1f MB (written in %.1f s)" % (args.records, os.path.getsize(p) / 1e6, time.perf_counter() - t0))
    for threads in (1, 0):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            from rectorch_b200.data import _Csv
            f = _Csv(p, ",", threads)
            m = f.to_csr(0, args.users, args.items, False)
            f.close()
            best = min(best, time.perf_counter() - t0)
        print("b200vae ingest, %s threads: %.3f s  (%.1f M records/s, %.0f MB/s)  nnz %d" % (
            "all" if threads == 0 else threads, best, args.records / best / 1e6, os.path.getsize(p) / best / 1e6, m.nnz))
    try:
        import pandas as pd
        from scipy import sparse
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            data = pd.read_csv(p)
            rows, cols = data['uid'], data['iid']
            ref = sparse.csr_matrix((np.ones_like(rows), (rows, cols)), dtype='float64', shape=(args.users, args.items))
            best = min(best, time.perf_counter() - t0)
        print("pandas.read_csv + scipy csr_matrix (reference path): %.3f s  (%.1f M records/s)  nnz %d" % (
            best, args.records / best / 1e6, ref.nnz))
        ref.sum_duplicates()
        ref.sort_indices()
        assert np.array_equal(ref.indptr, m.indptr) and np.array_equal(ref.indices, m.indices) and np.array_equal(ref.data, m.data)
        print("matrices identical")
    except ImportError:
        print("pandas not available")
