from .plans import OFPlan, ReducePlan, NxMPlan
from .ofbase import OFBaseBatch
from .algorithms import FeatureExtractors
from .filterdata import FilterData
from .plans import PSDPlan
from .noise import NoisePSD
from .oftrigger import OptimumFilterTrigger, TriggerPlan
from .eventbuilder import EventBuilder
