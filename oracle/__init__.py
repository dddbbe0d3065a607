"""CPU oracle for the xview inference + fusion + score() hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the product
package (``modular_semantic_segmentation_b200`` / ``xview``).  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg.

What it is: a literal CPU restatement (torch-CPU fp32 for the convolutions, numpy for
everything per-pixel, float64 twins where the reference runs float64 on the host) of the
reference graph, every function citing the reference file:line it follows
(paths relative to the reference repository root).

Pinning status:
  * score() measures  -- PINNED: reproduces the stored float64 measures of sacred run 868
    (``Experimental Details.ipynb`` cell 12) exactly, see tests/golden/exp868.npz.
  * Bayes fusion / decision matrix -- PINNED on self-consistency of the two reference
    statements (bayes_mix.py:12-58 vs :61-112) on the real 12x12 matrices of run 868.
  * Dirichlet parameter fit -- PINNED against outputs of the reference's own
    ``xview/models/dirichletDifferentiation.py`` / ``dirichlet_fastfit.py`` executed in the
    build container (tests/golden/make_golden.py, tests/golden/dirichlet_fit.npz).
  * bilinear kernel -- PINNED on the closed form custom_layers.py:13-21.
  * npz key layout -- PINNED on the 34-name list printed in
    ``Synthia Rand Cityscapes Examples.ipynb`` (tests/golden/fcn_weight_keys.json).
  * Adapnet forward numerics (oracle/adapnet.py) -- **PARITY UNPINNED** for the same reason;
    TF 'SAME' padding of the strided / atrous convolutions follows its documented rule and
    is checked on hand-computed cases (tests/test_oracle_golden.py).
  * FCN forward numerics (conv / pool / transposed conv / softmax at the TensorFlow
    boundary) -- **PARITY UNPINNED**: TensorFlow 1.x is not installable in the build image
    and the reference's own tests assert nothing numeric.  The restatement follows the
    documented TF semantics listed in SURVEY.md Appendix A.
"""

from .fcn import (bilinear_kernel_1d, bilinear_filter, glorot_fcn_params, fcn_param_shapes,
                  conv2d, deconv2d, max_pool2x2, dropout, encoder, decoder, fcn,
                  softmax, argmax_first, test_pipeline, cross_entropy, vgg16_tower,
                  fusion_fcn, fusion_fcn_params)
from .adapnet import (adapnet, adapnet_params, adapnet_param_shapes, same_padding, conv_bn,
                      block_a, block_b)
from .fusion import (bayes_conditionals, bayes_prior, bayes_fusion, bayes_decision_matrix,
                     dirichlet_log_norm, dirichlet_fusion, dirichlet_fusion_f32, dirichlet_prior,
                     dirichlet_uncertainty_fusion,
                     average_fusion, variance_fusion, mc_moments, normed_entropy,
                     sampling_uncertainty, sufficient_statistics)
from .score import confusion_matrix, score_measures
from .dirichlet_fit import (find_dirichlet_priors, fit_sufficient_statistic, init_a_moments,
                            ipsi, fixedpoint_fit)
