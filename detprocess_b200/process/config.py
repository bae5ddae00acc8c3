"""
``YamlConfig``: the YAML-driven processing configuration, with the semantics of the
reference's ``detprocess/process/config.py`` (feature section in full; trigger / salting /
didv / noise / template sections are carried through as ``{'overall', 'channels'}`` maps
because only the feature path is in scope):

* duplicate keys are an error (reference :666-684), ``include:`` files are merged (:131-138)
* obsolete keys are renamed (:71-79): nb_samples -> trace_length_samples, psd_tag/noise_tag ->
  csd_tag, nb_pretrigger_samples -> pretrigger_length_samples, ...
* top-level ``filter_file`` / ``didv_file`` are global; a top-level ``global:`` block is the
  feature "overall" block; unknown top-level keys are feature channels (:205-210)
* channel key ``all`` and comma lists expand; ``disable: True`` / ``run: False`` drop a channel
* every algorithm block needs ``run``; ``run: False`` blocks are removed
* trace lengths resolve algorithm > channel > global, samples or msec (:452-577)
* outputs ``channels``, ``channel_list``, ``traces_config {(N, P): [chans]}``, ``weights``, ``overall``
"""
import copy

import yaml
from yaml.loader import SafeLoader

from ..utils import utils

__all__ = ['YamlConfig']

FIELDS = ['salting', 'feature', 'didv', 'noise', 'template', 'trigger']
OVERALL = {
    'global': ['filter_file', 'didv_file'],
    'trigger': ['coincident_window_msec', 'coincident_window_samples'],
    'salting': ['dm_pdf_file', 'coincident_salts', 'energies', 'nsalt', 'do_salt_deadtime'],
    'feature': ['trace_length_samples', 'pretrigger_length_samples', 'trace_length_msec',
                'pretrigger_length_msec'],
}
OBSOLETE = {
    'trigger_name': 'trigger_channel',
    'nb_samples': 'trace_length_samples',
    'nb_pretrigger_samples': 'pretrigger_length_samples',
    'template_time_tags': 'template_group_ids',
    'psd_tag': 'csd_tag',
    'noise_tag': 'csd_tag',
    'deadtime_salt': 'do_salt_deadtime',
}


class _UniqueKeyLoader(SafeLoader):
    def construct_mapping(self, node, deep=False):
        if not isinstance(node, yaml.MappingNode):
            raise yaml.constructor.ConstructorError(
                None, None, 'expected a mapping node, but found %s' % node.id, node.start_mark)
        out = {}
        for key_node, value_node in node.value:
            key = self.construct_object(key_node, deep=deep)
            if key in out:
                raise ValueError(f'ERROR: Duplicate key "{key}" found in the yaml file for same channel '
                                 f'and algorithm. This is not allowed to avoid unwanted configuration!')
            out[key] = self.construct_object(value_node, deep=deep)
        return out


def _rename(d, old, new):
    if not isinstance(d, dict):
        return d
    for key in list(d.keys()):
        if isinstance(d[key], dict):
            _rename(d[key], old, new)
        if key == old:
            d[new] = d.pop(old)
    return d


class YamlConfig:
    def __init__(self, yaml_file, available_channels, sample_rate=None, verbose=True):
        self._yaml_file = yaml_file
        self._sample_rate = sample_rate
        if isinstance(available_channels, str):
            available_channels = [available_channels]
        self._available_channels = available_channels
        self._processing_config = None
        self._read_config()

    def get_config(self, processing_type=None):
        if self._processing_config is None:
            return None
        if processing_type is None:
            return copy.deepcopy(self._processing_config)
        if processing_type not in FIELDS:
            raise ValueError(f'ERROR: Configuration type "{processing_type}" not found!')
        return copy.deepcopy(self._processing_config[processing_type])

    # ------------------------------------------------------------------
    def _load(self, path):
        with open(path, 'r') as f:
            return yaml.load(f, Loader=_UniqueKeyLoader)

    def _read_config(self):
        if isinstance(self._yaml_file, dict):
            ydict = copy.deepcopy(self._yaml_file)
        else:
            ydict = self._load(self._yaml_file)
        if not ydict:
            raise ValueError('ERROR: No configuration loadedSomething went wrong...')
        if 'include' in ydict:
            inc = ydict.pop('include')
            for f in ([inc] if isinstance(inc, str) else inc):
                ydict.update(self._load(f))
        for old, new in OBSOLETE.items():
            ydict = _rename(ydict, old, new)

        cfg = {'global': {}}
        for field in FIELDS:
            cfg[field] = {'overall': {}, 'channels': {}}
        for p in OVERALL['global']:
            cfg['global'][p] = copy.deepcopy(ydict.pop(p)) if p in ydict else None

        for field in FIELDS:
            if field not in ydict:
                continue
            fmap = {'overall': {}, 'channels': {}}
            overall = OVERALL.get(field, [])
            for key, items in copy.deepcopy(ydict.pop(field)).items():
                if key in overall:
                    fmap['overall'][key] = items
                elif field == 'feature' and key == 'global':
                    for p in items:
                        fmap['overall'][p] = items[p]
                else:
                    fmap['channels'][key] = items
            cfg[field] = fmap
        # whatever is left at top level belongs to the feature section
        for key in ydict:
            if key == 'global':
                cfg['feature']['overall'] = copy.deepcopy(ydict[key])
            else:
                cfg['feature']['channels'][key] = copy.deepcopy(ydict[key])

        # expand "all" and comma lists, drop disabled channels
        for field in FIELDS:
            new = {}
            for chan, cdict in cfg[field]['channels'].items():
                if isinstance(cdict, dict) and (cdict.get('disable') or ('run' in cdict and not cdict['run'])):
                    continue
                if chan == 'all':
                    for c in self._available_channels:
                        new[c] = copy.deepcopy(cdict)
                else:
                    parts, _ = utils.split_channel_name(chan, available_channels=self._available_channels,
                                                        separator=',', label=field)
                    for c in parts:
                        new[c] = copy.deepcopy(cdict)
            cfg[field]['channels'] = new

        cfg['feature'] = self._configure_features(cfg['feature'], cfg['global'])
        cfg['trigger'] = self._configure_triggers(cfg['trigger'], cfg['global'])
        cfg['salting'] = self._configure_simple(cfg['salting'], cfg['global'], 'salting')
        self._processing_config = cfg

    def _configure_simple(self, section, global_config, label):
        out = copy.deepcopy(section)
        for k, v in (global_config or {}).items():
            out['overall'].setdefault(k, v)
        chans = []
        for chan, cc in out['channels'].items():
            if not isinstance(cc, dict):
                raise ValueError(f'ERROR: Channel {chan} has no configuration! Remove from yaml file or disable it!')
            parts, _ = utils.split_channel_name(chan, available_channels=self._available_channels, label=label)
            chans.extend(parts)
        out['channel_list'] = utils.unique_list(chans)
        return out

    def _configure_triggers(self, trigger_config, global_config):
        """Trigger section (reference config.py:324-408): one entry per trigger -- the channel itself when the block has
        a 'run' key, else '<algorithm>_<trigger channel>' per algorithm sub-block; 'trigger_channel' renames the
        channel; every entry carries the raw 'channel_name'."""
        out = copy.deepcopy(trigger_config)
        for k, v in (global_config or {}).items():
            out['overall'].setdefault(k, v)
        split, entries = [], {}
        for chan, cc in trigger_config['channels'].items():
            cc = copy.deepcopy(cc)
            if not isinstance(cc, dict):
                raise ValueError(f'ERROR: Channel {chan} has no configuration! Remove from yaml file or disable it!')
            parts, _ = utils.split_channel_name(chan, available_channels=self._available_channels, label='trigger')
            split.extend(parts)
            trigger_channel = cc.pop('trigger_channel', chan)
            if 'run' in cc:
                if not cc['run']:
                    continue
                cc['channel_name'] = chan
                entries[trigger_channel] = cc
            else:
                for algo, ad in cc.items():
                    if not isinstance(ad, dict) or 'run' not in ad:
                        raise ValueError(f'ERROR: Missing "run" parameter for trigger channel {chan}')
                    if not ad['run']:
                        continue
                    ad['channel_name'] = chan
                    entries[f'{algo}_{trigger_channel}'] = ad
        out['channels'] = entries
        out['channel_list'] = utils.unique_list(split)
        return out

    def _length(self, cfg, kind, default):
        """samples key wins over msec key; returns default when neither is present"""
        if f'{kind}_length_samples' in cfg:
            return cfg[f'{kind}_length_samples']
        if f'{kind}_length_msec' in cfg:
            if self._sample_rate is None:
                raise ValueError('ERROR: sample rate is required when trace length is in msec. ')
            return utils.convert_length_msec_to_samples(cfg[f'{kind}_length_msec'], self._sample_rate)
        return default

    def _configure_features(self, feature_config, global_config):
        fd = copy.deepcopy(feature_config)
        for k, v in (global_config or {}).items():
            fd['overall'].setdefault(k, v)
        split_all = []
        for chan in list(fd['channels'].keys()):
            cc = copy.deepcopy(fd['channels'][chan])
            if not isinstance(cc, dict):
                raise ValueError(f'ERROR: Channel {chan} has no configuration! Remove from yaml file or disable it!')
            parts, _ = utils.split_channel_name(chan, self._available_channels, label='feature')
            split_all.extend(parts)
            nb = self._length(cc, 'trace', self._length(fd['overall'], 'trace', None))
            npre = self._length(cc, 'pretrigger', self._length(fd['overall'], 'pretrigger', None))
            if nb is not None and npre is None:
                raise ValueError(f'ERROR: Missing "pretrigger_length_samples" for channel {chan} !')
            if nb is None and npre is not None:
                raise ValueError(f'ERROR: Missing "trace_length_samples"  for channel {chan} !')
            algos = []
            for algo, ac in cc.items():
                if not isinstance(ac, dict):
                    continue
                if 'run' not in ac:
                    raise ValueError(f'ERROR: Missing "run" parameter for channel {chan}, algorithm {algo}. '
                                     f'Please fix the configuration yaml file')
                if not ac['run']:
                    fd['channels'][chan].pop(algo)
                    continue
                algos.append(algo)
                fd['channels'][chan][algo]['nb_samples'] = self._length(ac, 'trace', nb)
                fd['channels'][chan][algo]['nb_pretrigger_samples'] = self._length(ac, 'pretrigger', npre)
            if not algos:
                fd['channels'].pop(chan)
            else:
                fd['channels'][chan].pop('trace_length_samples', None)
                fd['channels'][chan].pop('pretrigger_length_samples', None)
        fd['channel_list'] = utils.unique_list(split_all)

        traces_config, weights = {}, {}
        for chan, cc in fd['channels'].items():
            parts, _ = utils.split_channel_name(chan, fd['channel_list'])
            for c in parts:
                if f'weight_{c}' in cc:
                    weights.setdefault(chan, {})[f'weight_{c}'] = cc[f'weight_{c}']
            for algo, ac in cc.items():
                if not isinstance(ac, dict) or not ac['run']:
                    continue
                key = (ac['nb_samples'], ac['nb_pretrigger_samples'])
                traces_config.setdefault(key, []).extend(list(parts))
        for key in traces_config:
            traces_config[key] = utils.unique_list(traces_config[key])
        fd['traces_config'] = copy.deepcopy(traces_config) if traces_config else None
        fd['weights'] = copy.deepcopy(weights)
        return fd
