from .config import YamlConfig
from .features import FeatureProcessing
from .triggers import TriggerProcessing
