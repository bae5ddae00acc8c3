"""
Feature-table output in dumps, the way the reference's event loop does it (``detprocess/process/features.py:584-629``):
rows accumulate in memory and are written as ``<prefix>_F0001``, ``_F0002``, ... whenever the memory limit is reached
and at the end.  The reference exports vaex-HDF5; vaex / HDF5 are not available here, so the dumps are parquet files
(pyarrow) with the same columns and the same file-name scheme.
"""
import os

__all__ = ['FeatureWriter']


class FeatureWriter:
    def __init__(self, save_path, prefix='feature', series_name=None, memory_limit_gb=2.0):
        self._dir = save_path
        os.makedirs(save_path, exist_ok=True)
        self._prefix = prefix if series_name is None else f'{prefix}_{series_name}'
        self._limit = float(memory_limit_gb) * 1e9
        self._frames = []
        self._bytes = 0
        self._dump = 1
        self.files = []

    def add(self, rows):
        """queue a batch of rows (a DataFrame or a dict column name -> ndarray); dumps when the queued rows exceed the
        memory limit"""
        import numpy as np
        if rows is None or len(rows) == 0:
            return
        if isinstance(rows, dict):
            self._bytes += int(sum(np.asarray(v).nbytes for v in rows.values()))
        else:
            self._bytes += int(rows.memory_usage(index=False, deep=True).sum())
            rows = {c: rows[c].to_numpy() for c in rows.columns}
        self._frames.append(rows)
        if self._bytes >= self._limit:
            self.flush()

    def flush(self):
        from ..utils.utils import columns_to_frame
        if not self._frames:
            return None
        df = columns_to_frame(self._frames)
        name = os.path.join(self._dir, f'{self._prefix}_F{self._dump:04d}.parquet')
        df.to_parquet(name)
        self.files.append(name)
        self._dump += 1
        self._frames, self._bytes = [], 0
        return name

    def close(self):
        self.flush()
        return self.files
