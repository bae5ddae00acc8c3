"""
Regenerates the YamlConfig golden fixture from the reference's own artefacts (run in the
build container, where /root/reference exists):

  input : /root/reference/examples/processing/process_example.yaml      (parsed -> JSON)
  output: the saved `pprint(YamlConfig(...).get_config())` of
          /root/reference/examples/processing/test_reading_yaml.ipynb cell 6
          (the only pinned result anywhere in the reference)

Only the 'feature' and 'global' sections are kept: the notebook's trigger section has
drifted from the YAML.
"""
import ast
import json
import os

import yaml

REF = '/root/reference/examples/processing'
HERE = os.path.dirname(os.path.abspath(__file__))

nb = json.load(open(os.path.join(REF, 'test_reading_yaml.ipynb')))
gold = ast.literal_eval(''.join(nb['cells'][6]['outputs'][0]['text']))
with open(os.path.join(HERE, 'yaml_config_expected.pyl'), 'w') as f:
    f.write(repr({'feature': gold['feature'], 'global': gold['global']}))
parsed = yaml.safe_load(open(os.path.join(REF, 'process_example.yaml')))
with open(os.path.join(HERE, 'yaml_config_input.json'), 'w') as f:
    json.dump({'yaml': parsed, 'sample_rate': 1.25e6,
               'available_channels': ['Melange025pcLeft', 'Melange025pcRight', 'Melange4pc1ch', 'Melange1pc1ch']},
              f, indent=1)
print('golden written')
