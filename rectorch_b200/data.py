"""Reading of pre-processed data sets (rectorch/data.py:328-409, 499-560) without pandas.

``DataReader(data_config).load_data(datatype)`` and ``DatasetManager(config)`` keep the reference's
signatures, return types (``scipy.sparse.csr_matrix`` of float64, canonical form) and error behaviour; the
CSV files are parsed and turned into CSR by the multi-threaded host ingest of libb200vae.so
(``b200vae_csv_*``, csrc/ingest.cu) instead of ``pd.read_csv`` + ``csr_matrix((values, (rows, cols)))``.
The matrices go straight into :class:`rectorch_b200.samplers.DataSampler`, which uploads them to HBM once.

Not mirrored (host-side ETL outside the training path): ``DataProcessing`` (raw csv ->
pre-processed files) and ``DataReader.load_data_as_dict`` (sequence view used by SVAE only).
"""
import ctypes
import os

import numpy as np
from scipy import sparse

from . import _lib
from ._lib import check
from .configuration import DataConfig

__all__ = ['DataReader', 'DatasetManager', 'read_csv_csr']


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
