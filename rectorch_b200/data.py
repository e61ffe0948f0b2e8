"""Reading of pre-processed data sets (rectorch/data.py:328-409, 499-560) without pandas.

``DataReader(data_config).load_data(datatype)`` and ``DatasetManager(config)`` keep the reference's
signatures, return types (``scipy.sparse.csr_matrix`` of float64, canonical form) and error behaviour; the
CSV files are parsed and turned into CSR by the multi-threaded host ingest of libb200vae.so
(``b200vae_csv_*``, csrc/ingest.cu) instead of ``pd.read_csv`` + ``csr_matrix((values, (rows, cols)))``.
The matrices go straight into :class:`rectorch_b200.samplers.DataSampler`, which uploads them to HBM once.

``DataProcessing(data_config).process()`` (data.py:46-325) is restated on numpy arrays: the reference's version
depends on pandas group-by semantics that current pandas releases no longer have (``groupby(..., as_index=False)
.size()`` returning a Series), so it cannot run in this environment; the restatement follows its steps and its
``numpy.random`` call sequence one by one and is pinned by the exact file contents the reference's own tests expect
(tests/test_data.py:14-101, 280-350).  Only the two ends use pandas, when it is importable: ``pd.read_csv`` for the raw
file and ``DataFrame.to_csv`` for the output (what the reference itself calls); a plain-Python path writes
byte-identical files without it.

Not mirrored: ``DataReader.load_data_as_dict`` (sequence view used by SVAE only).
"""
import logging
import ctypes
import os

import numpy as np
from scipy import sparse

from . import _lib
from ._lib import check
from .configuration import DataConfig

__all__ = ['DataProcessing', 'DataReader', 'DatasetManager', 'read_csv_csr']

logger = logging.getLogger(__name__)


class _Csv:
    """One parsed rating file (records stay inside the library until converted)."""

    def __init__(self, path, sep=",", n_threads=0):
        self._h = ctypes.c_void_p()
        check(_lib.lib().b200vae_csv_open(ctypes.byref(self._h), os.fsencode(path), sep.encode()[:1], int(n_threads)))
        n, nc = ctypes.c_int64(), ctypes.c_int32()
        umin, umax, imax = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(_lib.lib().b200vae_csv_info(self._h, ctypes.byref(n), ctypes.byref(nc), ctypes.byref(umin),
                                          ctypes.byref(umax), ctypes.byref(imax)))
        self.n_records, self.n_cols = n.value, nc.value
        self.uid_min, self.uid_max, self.iid_max = umin.value, umax.value, imax.value
        self.value_column = _lib.lib().b200vae_csv_value_column(self._h).decode()

    def to_csr(self, uid_base, n_rows, n_items, use_values):
        indptr = np.empty(n_rows + 1, dtype=np.int64)
        indices = np.empty(max(self.n_records, 1), dtype=np.int32)
        values = np.empty(max(self.n_records, 1), dtype=np.float64)
        nnz = ctypes.c_int64()
        check(_lib.lib().b200vae_csv_to_csr(self._h, int(uid_base), int(n_rows), int(n_items), 1 if use_values else 0,
                                            indptr.ctypes.data_as(ctypes.c_void_p),
                                            indices.ctypes.data_as(ctypes.c_void_p),
                                            values.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nnz)))
        k = nnz.value
        m = sparse.csr_matrix((values[:k].copy(), indices[:k].copy(), indptr), shape=(n_rows, n_items))
        m.has_sorted_indices = True
        return m

    def close(self):
        if self._h:
            _lib.lib().b200vae_csv_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_csv_csr(path, n_items, topn=True, uid_base=None, n_rows=None, sep=","):
    """``uid,iid[,value]`` file -> canonical float64 CSR of shape ``[n_rows x n_items]`` (row = uid - uid_base;
    defaults: uid_base 0, n_rows = max uid + 1)."""
    f = _Csv(path, sep)
    try:
        base = 0 if uid_base is None else uid_base
        rows = (f.uid_max - base + 1) if n_rows is None else n_rows
        return f.to_csr(base, max(rows, 0), n_items, not topn)
    finally:
        f.close()


def _as_column(values):
    """pandas-style inference of one raw column: int64 if every field is an integer literal, else float64 if every
    field parses as a float, else the strings themselves."""
    try:
        return np.array([int(v) for v in values], dtype=np.int64)
    except ValueError:
        pass
    try:
        return np.array([float(v) for v in values], dtype=np.float64)
    except ValueError:
        return np.array(values, dtype=object)


def _fmt(v):
    """Field formatting of DataFrame.to_csv / '%s' for the inferred column types."""
    if isinstance(v, (np.floating, float)):
        return repr(float(v))
    return str(v)


class DataProcessing:
    """Pre-processing of a raw rating file into the ``proc_path`` folder (rectorch/data.py:46-325).

    ``DataProcessing(data_config)`` takes a :class:`DataConfig` or the path of its JSON file (anything else raises
    :class:`TypeError`); :meth:`process` runs the reference's pipeline: read the csv, keep ratings above
    ``threshold``, drop items / users with fewer than ``i_min`` / ``u_min`` ratings, permute the users
    (``np.random.permutation`` under ``seed``) and split them into training / validation / test users
    (``heldout`` each for the last two), restrict validation / test ratings to training items, drop held-out users
    with fewer than two ratings, split each held-out user's ratings into a training and a test part
    (``test_prop``, at least one test item, ``np.random.choice`` under a re-seeded generator), build the id maps
    (``u2id`` / ``i2id``) and write ``train.csv``, ``validation_{tr,te}.csv``, ``test_{tr,te}.csv``,
    ``unique_uid.txt`` and ``unique_iid.txt``.
    """

    def __init__(self, data_config):
        if isinstance(data_config, DataConfig):
            self.cfg = data_config
        elif isinstance(data_config, str):
            self.cfg = DataConfig(data_config)
        else:
            raise TypeError("'data_config' must be of type 'DataConfig' or 'str'.")
        self.i2id = {}
        self.u2id = {}

    # -- raw file ------------------------------------------------------------------------------------------
    def _read_raw(self):
        """(column names, one numpy array per column).  pandas' C parser does the work when pandas is importable
        (it is what the reference calls, data.py:101-104: same type inference, seconds instead of minutes on
        ml-20m); `_read_raw_py` is the dependency-free equivalent."""
        try:
            import pandas as pd
        except Exception:                            # pragma: no cover - pandas is optional
            return self._read_raw_py()
        sep = self.cfg.separator if self.cfg.separator else ','
        header = None if self.cfg.header is None else int(self.cfg.header)
        try:
            df = pd.read_csv(self.cfg.data_path, sep=sep, header=header, engine='c' if len(sep) == 1 else 'python')
        except Exception:
            return self._read_raw_py()
        names = [c.strip() if isinstance(c, str) else int(c) for c in df.columns]
        cols = []
        for c in df.columns:
            col = df[c]
            if col.dtype.kind in "iu":
                cols.append(col.to_numpy(dtype=np.int64))
            elif col.dtype.kind == "f":
                cols.append(col.to_numpy(dtype=np.float64))
            else:
                cols.append(np.array([v.strip() if isinstance(v, str) else v for v in col.tolist()], dtype=object))
        return names, cols

    def _read_raw_py(self):
        sep = self.cfg.separator if self.cfg.separator else ','
        with open(self.cfg.data_path, "r") as fh:
            lines = [ln.rstrip("\r\n") for ln in fh if ln.strip()]
        names = None
        if self.cfg.header is not None:              # pandas: header=<row number of the column names>
            h = int(self.cfg.header)
            names = lines[h].split(sep)
            lines = lines[h + 1:]
        fields = [ln.split(sep) for ln in lines]
        n_cols = len(fields[0]) if fields else 0
        cols = [_as_column([f[c].strip() for f in fields]) for c in range(n_cols)]
        if names is None:
            names = list(range(n_cols))             # header=None: pandas names the columns 0, 1, 2, ...
        return names, cols

    @staticmethod
    def _counts(keys):
        """group-by size: (sorted unique keys, count of each)."""
        return np.unique(keys, return_counts=True)

    def _split_train_test(self, uid, rows):
        """data.py:290-312 on row indices: per held-out user (ascending id) mark max(int(test_prop * n), 1) of the
        n ratings as test items with np.random.choice (generator re-seeded with cfg.seed)."""
        np.random.seed(self.cfg.seed)
        test_prop = float(self.cfg.test_prop) if self.cfg.test_prop else 0.2
        tr_list, te_list = [], []
        # one stable sort groups the rows by user (ascending id, file order kept inside a group): the same groups,
        # in the same order, as a boolean mask per user -- and so the same np.random.choice sequence
        u_of = uid[rows]
        order = np.argsort(u_of, kind='stable')
        _, starts = np.unique(u_of[order], return_index=True)
        for grp in np.split(rows[order], starts[1:]) if len(rows) else []:
            n_items_u = len(grp)
            if n_items_u > 1:
                idx = np.zeros(n_items_u, dtype='bool')
                sz = max(int(test_prop * n_items_u), 1)
                idx[np.random.choice(n_items_u, size=sz, replace=False).astype('int64')] = True
                tr_list.append(grp[np.logical_not(idx)])
                te_list.append(grp[idx])
            else:
                logger.warning("Skipped user in test set: number of ratings <= 1.")
        cat = lambda lst: np.concatenate(lst) if lst else np.zeros(0, dtype=np.int64)   # noqa: E731
        return cat(tr_list), cat(te_list)

    def process(self):
        """Run the whole pre-processing (data.py:98-213)."""
        np.random.seed(int(self.cfg.seed))
        logger.info("Reading data file %s.", self.cfg.data_path)
        names, cols = self._read_raw()
        keep = np.arange(len(cols[0]) if cols else 0)
        if self.cfg.threshold:
            keep = keep[cols[2][keep] > float(self.cfg.threshold)]
        logger.info("Applying filtering.")
        uid, iid = cols[0], cols[1]
        imin, umin = int(self.cfg.i_min), int(self.cfg.u_min)
        if imin > 0:
            items, cnt = self._counts(iid[keep])
            keep = keep[np.isin(iid[keep], items[cnt >= imin])]
        if umin > 0:
            users, cnt = self._counts(uid[keep])
            keep = keep[np.isin(uid[keep], users[cnt >= umin])]
        unique_uid = np.unique(uid[keep])                       # group-by index: ascending user ids
        idx_perm = np.random.permutation(unique_uid.size)
        unique_uid = unique_uid[idx_perm]
        n_users = unique_uid.size
        n_heldout = int(self.cfg.heldout)

        logger.info("Calculating splits.")
        tr_users = unique_uid[:(n_users - n_heldout * 2)]
        vd_users = unique_uid[(n_users - n_heldout * 2): (n_users - n_heldout)]
        te_users = unique_uid[(n_users - n_heldout):]
        train_rows = keep[np.isin(uid[keep], tr_users)]
        _, first = np.unique(iid[train_rows], return_index=True)
        unique_iid = iid[train_rows][np.sort(first)]            # pd.unique: order of first appearance

        logger.info("Creating validation and test set.")

        def heldout_rows(users):
            rows = keep[np.isin(uid[keep], users)]
            rows = rows[np.isin(iid[rows], unique_iid)]
            us, cnt = self._counts(uid[rows])
            kept = rows[np.isin(uid[rows], us[cnt >= 2])]
            return kept, len(us) - len(np.unique(uid[kept]))

        val_rows, vdiff = heldout_rows(vd_users)
        test_rows, tdiff = heldout_rows(te_users)
        if vdiff > 0:
            logger.warning("Skipped %d users in validation set.", vdiff)
        if tdiff > 0:
            logger.warning("Skipped %d users in test set.", tdiff)
        val_tr, val_te = self._split_train_test(uid, val_rows)
        test_tr, test_te = self._split_train_test(uid, test_rows)

        us = set(np.unique(uid[val_rows]).tolist()) | set(np.unique(uid[test_rows]).tolist())
        unique_uid = list(unique_uid.tolist())
        unique_uid = unique_uid[:len(tr_users)] + [u for u in unique_uid[len(tr_users):] if u in us]
        self.i2id = dict((i, k) for (k, i) in enumerate(unique_iid.tolist()))
        self.u2id = dict((u, k) for (k, u) in enumerate(unique_uid))

        pro_dir = self.cfg.proc_path
        if not os.path.exists(pro_dir):
            os.makedirs(pro_dir)
        logger.info("Saving unique_iid.txt.")
        with open(os.path.join(pro_dir, 'unique_iid.txt'), 'w') as f:
            for i in unique_iid.tolist():
                f.write('%s\n' % _fmt(i))
        logger.info("Saving unique_uid.txt.")
        with open(os.path.join(pro_dir, 'unique_uid.txt'), 'w') as f:
            for u in unique_uid:
                f.write('%s\n' % _fmt(u))

        logger.info("Saving all the files.")
        extra = [] if self.cfg.topn else list(range(2, len(cols)))

        def mapped(keys, table):
            """table[key] for every key: one sorted lookup for integer ids, a dict pass for string ids."""
            if keys.dtype.kind in "iu" and table:
                ks = np.fromiter(table.keys(), dtype=np.int64, count=len(table))
                vs = np.fromiter(table.values(), dtype=np.int64, count=len(table))
                order = np.argsort(ks)
                return vs[order][np.searchsorted(ks[order], keys)]
            return np.array([table[k] for k in keys.tolist()], dtype=np.int64)

        def save(name, rows):
            u, i = mapped(uid[rows], self.u2id), mapped(iid[rows], self.i2id)
            head = ['uid', 'iid'] + [str(names[c]) for c in extra]
            path = os.path.join(pro_dir, name)
            try:                                     # DataFrame.to_csv is what the reference writes with
                import pandas as pd
                data = {'uid': u, 'iid': i}
                for c in extra:
                    data[str(names[c])] = cols[c][rows]
                pd.DataFrame(data, columns=head).to_csv(path, index=False)
                return
            except ImportError:                      # pragma: no cover - pandas is optional
                pass
            fields = [u.tolist(), i.tolist()] + [[_fmt(v) for v in cols[c][rows].tolist()] for c in extra]
            with open(path, 'w') as f:
                f.write(",".join(head) + "\n")
                f.write("".join(",".join(map(str, rec)) + "\n" for rec in zip(*fields)))

        save('train.csv', train_rows)
        save('validation_tr.csv', val_tr)
        save('validation_te.csv', val_te)
        save('test_tr.csv', test_tr)
        save('test_te.csv', test_te)
        logger.info("Preprocessing complete!")


class DataReader():
    """Utility class for reading a pre-processed data set (rectorch/data.py:328-409).

    Parameters
    ----------
    data_config : :class:`rectorch_b200.configuration.DataConfig` or :obj:`str`
        The data configuration or the path of its JSON file; anything else raises :class:`TypeError`.

    Attributes
    ----------
    cfg : the configuration;  n_items : number of lines of ``unique_iid.txt``.
    """

    def __init__(self, data_config):
        if isinstance(data_config, DataConfig):
            self.cfg = data_config
        elif isinstance(data_config, str):
            self.cfg = DataConfig(data_config)
        else:
            raise TypeError("'data_config' must be of type 'DataConfig' or 'str'.")
        self.n_items = self._load_n_items()

    def load_data(self, datatype='train'):
        """``'train'`` / ``'full'`` -> one csr_matrix; ``'validation'`` / ``'test'`` -> (tr, te) pair
        (data.py:363-373); any other string raises :class:`ValueError`."""
        if datatype == 'train':
            return self._load_train_data()
        elif datatype == 'validation':
            return self._load_train_test_data(datatype)
        elif datatype == 'test':
            return self._load_train_test_data(datatype)
        elif datatype == 'full':
            tr = self._load_train_data()
            val_tr, val_te = self._load_train_test_data("validation")
            te_tr, te_te = self._load_train_test_data("test")
            val = val_tr + val_te
            te = te_tr + te_te
            return sparse.vstack([tr, val, te])
        else:
            raise ValueError("Possible datatype values are 'train', 'validation', 'test', 'full'.")

    def _load_n_items(self):
        n = 0
        with open(os.path.join(self.cfg.proc_path, 'unique_iid.txt'), 'r') as f:
            for _ in f:
                n += 1
        return n

    def _load_train_data(self):
        # data.py:375-391: n_users = max uid + 1, ones when cfg.topn else the third column
        path = os.path.join(self.cfg.proc_path, 'train.csv')
        return read_csv_csr(path, self.n_items, topn=bool(self.cfg.topn))

    def _load_train_test_data(self, datatype='test'):
        # data.py:393-420: rows are uid - min uid over both files; users without a training item are dropped
        tr_path = os.path.join(self.cfg.proc_path, '%s_tr.csv' % datatype)
        te_path = os.path.join(self.cfg.proc_path, '%s_te.csv' % datatype)
        f_tr, f_te = _Csv(tr_path), _Csv(te_path)
        try:
            start_idx = min(f_tr.uid_min, f_te.uid_min)
            end_idx = max(f_tr.uid_max, f_te.uid_max)
            n_rows = end_idx - start_idx + 1
            use_values = not bool(self.cfg.topn)
            data_tr = f_tr.to_csr(start_idx, n_rows, self.n_items, use_values)
            data_te = f_te.to_csr(start_idx, n_rows, self.n_items, use_values)
        finally:
            f_tr.close()
            f_te.close()
        tr_idx = np.diff(data_tr.indptr) != 0
        return data_tr[tr_idx], data_te[tr_idx]


class DatasetManager():
    """Training / validation / test sets of one configuration (rectorch/data.py:499-560)."""

    def __init__(self, config_file):
        reader = DataReader(config_file)
        train_data = reader.load_data('train')
        vad_data_tr, vad_data_te = reader.load_data('validation')
        test_data_tr, test_data_te = reader.load_data('test')

        self.n_items = reader.n_items
        self.training_set = (train_data, None)
        self.validation_set = (vad_data_tr, vad_data_te)
        self.test_set = (test_data_tr, test_data_te)

    def get_train_and_test(self):
        """Training + validation + training part of the test users as one training matrix; the test part of
        the test users (last rows) as the test matrix (data.py:543-560)."""
        tr = sparse.vstack([self.training_set[0], sum(self.validation_set), self.test_set[0]])
        shape = tr.shape[0] - self.test_set[1].shape[0], tr.shape[1]
        te = sparse.vstack([sparse.csr_matrix(shape), self.test_set[1]])
        return tr, te
