from .config import YamlConfig
from .features import FeatureProcessing
