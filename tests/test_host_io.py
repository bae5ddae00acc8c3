"""CPU: the raw-event reader interface and the dumped feature writer (HDF5 is not available, the
pipeline talks to detprocess_b200.io.EventReader)."""
import json
import os

import numpy as np
import pandas as pd
import pytest

from detprocess_b200.io import ArrayReader, RawBinaryReader, FeatureWriter, write_raw_binary


def test_raw_binary_roundtrip_and_adc_conversion(tmp_path):
    rng = np.random.default_rng(1)
    adc = rng.integers(-30000, 30000, size=(37, 2, 512)).astype(np.int16)
    base = str(tmp_path / 'series_0001')
    write_raw_binary(base, adc, ['chanA', 'chanB'], 1.25e6, adc_gain=[2e-11, 3e-11], adc_offset=[1e-9, -2e-9],
                     admin={'event_number': np.arange(37) + 100, 'series_number': [7] * 37})
    r = RawBinaryReader(base)
    assert len(r) == 37 and r.channels == ['chanA', 'chanB'] and r.sample_rate == 1.25e6
    b = r.read_batch(5, 21, pinned=False)
    assert b.dtype.is_floating_point is False and tuple(b.shape) == (16, 2, 512)
    assert np.array_equal(b.numpy(), adc[5:21])
    b2 = r.read_batch(21, 37, pinned=False)                  # second staging buffer: the first batch is still intact
    assert np.array_equal(b.numpy(), adc[5:21]) and np.array_equal(b2.numpy(), adc[21:37])
    amps = r.to_amps(b).numpy()
    assert np.array_equal(amps[:, 1], adc[5:21, 1].astype(np.float64) * 3e-11 + -2e-9)
    assert np.array_equal(r.admin(5, 8)['event_number'], [105, 106, 107])
    with open(base + '.json') as f:
        meta = json.load(f)
    meta['n_events'] = 40
    with open(base + '.json', 'w') as f:
        json.dump(meta, f)
    with pytest.raises(ValueError):
        RawBinaryReader(base)                                # size does not match the metadata
    with pytest.raises(ValueError):
        write_raw_binary(base, adc.astype(np.int32), ['chanA', 'chanB'], 1.25e6)


def test_array_reader_wraps_the_dict_form():
    x = np.random.default_rng(2).standard_normal((9, 64))
    r = ArrayReader(x, ['only'], 1e6, admin={'event_number': np.arange(9) * 2})
    assert len(r) == 9 and tuple(r.read_batch(2, 5).shape) == (3, 1, 64)
    assert np.array_equal(r.to_amps(r.read_batch(0, 9)).numpy()[:, 0], x)
    assert np.array_equal(r.admin(1, 3)['event_number'], [2, 4])
    with pytest.raises(ValueError):
        ArrayReader(np.zeros((3, 2, 8)), ['a'], 1e6)


def test_feature_writer_dumps_by_memory_limit(tmp_path):
    w = FeatureWriter(str(tmp_path), prefix='feature', series_name='I2_D20230615_T231959', memory_limit_gb=1e-5)   # 10 kB
    frames = [pd.DataFrame({'event_number': np.arange(i * 300, (i + 1) * 300), 'amp': np.full(300, float(i))}) for i in range(5)]
    for df in frames:
        w.add(df)
    w.add(pd.DataFrame())
    files = w.close()
    assert [os.path.basename(f) for f in files][:2] == ['feature_I2_D20230615_T231959_F0001.parquet',
                                                        'feature_I2_D20230615_T231959_F0002.parquet']
    back = pd.concat([pd.read_parquet(f) for f in files], ignore_index=True)
    assert back.equals(pd.concat(frames, ignore_index=True))
    assert FeatureWriter(str(tmp_path / 'empty')).close() == []


def test_pytesio_reader_is_import_guarded():
    """the HDF5-backed EventReader subclass exists and fails loudly where pytesio is absent"""
    from detprocess_b200.io.h5 import PytesioReader
    from detprocess_b200.io import EventReader
    assert issubclass(PytesioReader, EventReader)
    try:
        import pytesio  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match='pytesio'):
            PytesioReader(['whatever.hdf5'])
