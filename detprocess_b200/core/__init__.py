from .plans import OFPlan, ReducePlan
from .ofbase import OFBaseBatch
from .algorithms import FeatureExtractors
from .filterdata import FilterData
