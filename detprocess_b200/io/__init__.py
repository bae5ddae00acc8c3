from .readers import EventReader, ArrayReader, RawBinaryReader, write_raw_binary
from .writers import FeatureWriter
