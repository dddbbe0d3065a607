"""`xview.models` facade over modular_semantic_segmentation_b200.models (same names as
xview/models/__init__.py:1-26 of the reference)."""
import sys

from modular_semantic_segmentation_b200 import models as _impl
from modular_semantic_segmentation_b200.models import (Adapnet, AverageFusion, BayesFusion,
                                                       BayesianFCN, DirichletFusion, FusionFCN,
                                                       SimpleFCN, VarianceFusion, get_model)

# make `xview.models.simple_fcn`, `xview.models.bayes_mix`, ... importable as in the reference
for _name in ('base_model', 'simple_fcn', 'basic_fusion_model', 'bayes_mix', 'dirichlet_mix',
              'average_mix', 'variance_mix', 'custom_layers', 'dirichletDifferentiation',
              'fusion_fcn', 'adapnet', 'bayesian_fcn', 'uncertainty_dirichlet_mix'):
    __import__('modular_semantic_segmentation_b200.models.' + _name)
    sys.modules[__name__ + '.' + _name] = getattr(_impl, _name)
    globals()[_name] = getattr(_impl, _name)
