"""
pytesio-backed ``EventReader``: the reference reads raw events with ``pytesio.H5Reader`` --
``read_many_events(output_format=2, ...)`` for batches (``core/noise.py:671-682``) and ``read_single_event(event_index,
trigger_index=..., trace_length_samples=..., pretrigger_length_samples=..., adctoamp=True)`` per event
(``process/processing_data.py:643-688``).  Neither pytesio nor an HDF5 library exists in this image, so this module is
IMPORT-GUARDED and untested here: constructing the reader without pytesio raises a clear error; with pytesio it hands
the pipeline what every other reader does -- events as stored (int16 ADC counts, ``adctoamp=False``) in pinned host
memory plus the per-channel linear ADC -> amps conversion in ``metadata``, so the conversion runs in the kernels' loads
and a quarter of the bytes crosses PCIe.  The vaex-HDF5 feature dumps of the reference (``features.py:595-616``) have the
same status: ``detprocess_b200.io.writers.FeatureWriter`` writes parquet with the reference's file-name scheme.
"""
import numpy as np

from .readers import EventReader, _pin

__all__ = ['PytesioReader']


class PytesioReader(EventReader):
    """``file_list``: raw pytesdaq HDF5 files of one series.  ``channels``: detector channels to read (file order kept)."""

    def __init__(self, file_list, channels=None):
        try:
            import pytesio
        except ImportError as e:      # the state of this image
            raise ImportError('PytesioReader needs pytesio (and HDF5), which this environment does not provide; use '
                              'RawBinaryReader / ArrayReader, or install pytesio next to detprocess_b200') from e
        self._h5 = pytesio.H5Reader()
        self._files = list(file_list)
        meta = self._h5.get_metadata(file_name=self._files[0], include_dataset_metadata=False)
        adc = meta['groups'][meta['adc_list'][0]]
        det = self._h5.get_detector_config(file_name=self._files[0])
        self._channels = list(channels) if channels is not None else list(det.keys())
        # per-event catalogue: (file index, event index) in file order
        self._catalogue = []
        for fi, f in enumerate(self._files):
            m = self._h5.get_metadata(file_name=f, include_dataset_metadata=False)
            n = int(m['groups'][m['adc_list'][0]]['nb_events'])
            self._catalogue.extend((fi, ei) for ei in range(n))
        conv = [self._adc_conversion(det[c], adc) for c in self._channels]
        self.metadata = {'sample_rate': float(adc['sample_rate']), 'channels': self._channels,
                         'nb_samples': int(adc['nb_samples']), 'dtype': 'int16',
                         'adc_gain': [g for g, _ in conv], 'adc_offset': [o for _, o in conv],
                         'detector_config': det, 'group_name': meta.get('group_name')}
        self._stage = [None, None]
        self._flip = 0

    @staticmethod
    def _adc_conversion(chan_cfg, adc_cfg):
        """amps = (adc * volts_per_count + volt_offset) / close_loop_norm: the linear map pytesio applies with adctoamp=True"""
        volts_per_count = float(adc_cfg['voltage_range'][1] - adc_cfg['voltage_range'][0]) / 2.0 ** 16
        offset_volts = float(adc_cfg['voltage_range'][1] + adc_cfg['voltage_range'][0]) / 2.0
        norm = float(chan_cfg['close_loop_norm'])
        return volts_per_count / norm, offset_volts / norm

    def __len__(self):
        return len(self._catalogue)

    def read_batch(self, i0, i1, pinned=True):
        import torch
        nb = i1 - i0
        k = self._flip
        self._flip ^= 1
        buf = self._stage[k]
        if buf is None or buf.shape[0] < nb:
            buf = _pin(torch.empty((nb, len(self._channels), self.metadata['nb_samples']), dtype=torch.int16), pinned)
            self._stage[k] = buf
        out = buf[:nb].numpy()
        for j, (fi, ei) in enumerate(self._catalogue[i0:i1]):
            out[j] = self._h5.read_single_event(ei, file_name=self._files[fi], detector_chans=self._channels, adctoamp=False)
        return buf[:nb]

    def admin(self, i0, i1):
        cols = {}
        for j, (fi, ei) in enumerate(self._catalogue[i0:i1]):
            info = self._h5.read_single_event(ei, file_name=self._files[fi], detector_chans=self._channels[:1],
                                              include_metadata=True, adctoamp=False)[1]['event']
            for key, val in info.items():
                cols.setdefault(key, []).append(val)
        return {k: np.asarray(v) for k, v in cols.items()}
