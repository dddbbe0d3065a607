"""B200-native counterparts of the `xview.models` classes (same names, same factory)."""
from .simple_fcn import SimpleFCN
from .bayes_mix import BayesFusion
from .dirichlet_mix import DirichletFusion
from .average_mix import AverageFusion
from .variance_mix import VarianceFusion
from .fusion_fcn import FusionFCN
from .adapnet import Adapnet
from .bayesian_fcn import BayesianFCN


def get_model(name):
    """xview/models/__init__.py:10-26."""
    if name == 'fcn':
        return SimpleFCN
    elif name == 'adapnet':
        return Adapnet
    elif name == 'bayesian_fcn':     # not in the reference's factory (its class cannot be built)
        return BayesianFCN
    elif name == 'fusion_fcn':
        return FusionFCN
    elif name in ['bayes_mix', 'bayes_fusion']:
        return BayesFusion
    elif name in ['dirichlet_mix', 'dirichlet_fusion']:
        return DirichletFusion
    elif name in ['average_fusion', 'average_mix']:
        return AverageFusion
    elif name in ['variance_mix', 'variance_fusion']:
        return VarianceFusion
    else:
        raise UserWarning('ERROR: Model %s not found' % name)
