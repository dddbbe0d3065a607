"""Generate the committed golden fixtures from the reference checkout.

Run ONCE in the build container (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py
Outputs (committed):
  exp868.npz            known-answer data of sacred run 868, parsed from the stored output of
                        `Experimental Details.ipynb` cell 12: per-expert measure-set confusion
                        matrices, fused test confusion matrix and the stored score() measures
  fcn_weight_keys.json  the variable names the reference prints for a bias-ful, BN-less FCN
                        expert (`Synthia Rand Cityscapes Examples.ipynb` rerun output)
  dirichlet_fit.npz     inputs/outputs of the reference's own host numerics
                        (xview/models/dirichletDifferentiation.py, dirichlet_fastfit.py)
                        executed here by file path (they only need numpy/scipy)
"""
import contextlib
import importlib.util
import io
import json
import os
import re

import numpy as np

REF = os.environ.get('XVIEW_REFERENCE', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def load_by_path(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_exp868():
    nb = json.load(open(os.path.join(REF, 'Experimental Details.ipynb')))
    cell = nb['cells'][12]
    text = None
    for out in cell['outputs']:
        if 'data' in out:
            text = ''.join(out['data']['text/plain'])
    info = eval(text, {'array': np.array, '__builtins__': {}})
    out = {
        'cm_measure_rgb': np.asarray(info['confusion_matrices']['rgb']),
        'cm_measure_depth': np.asarray(info['confusion_matrices']['depth']),
        'cm_fusion': np.asarray(info['confusion_matrix']),
    }
    for key in ('rgb', 'depth', 'fusion'):
        meas = info['measurements'][key]
        for name, value in meas.items():
            if name == 'confusion_matrix':
                if isinstance(value, dict):      # {'py/id': 2} = alias of info['confusion_matrix']
                    value = info['confusion_matrix']
                out['cm_test_%s' % key] = np.asarray(value)
            else:
                out['%s_%s' % (key, name)] = np.asarray(value)
    np.savez_compressed(os.path.join(HERE, 'exp868.npz'), **out)
    print('exp868.npz:', sorted(out))


def make_weight_keys():
    text = open(os.path.join(REF, 'Synthia Rand Cityscapes Examples.ipynb')).read()
    names = re.findall(r'WARNING: ((?:rgb|depth)/[a-z0-9_]+/(?:kernel|bias)) not found', text)
    keys = {'rgb': [], 'depth': []}
    for n in names:
        pre = n.split('/')[0]
        if n not in keys[pre]:
            keys[pre].append(n)
    json.dump(keys, open(os.path.join(HERE, 'fcn_weight_keys.json'), 'w'), indent=1)
    print('fcn_weight_keys.json:', {k: len(v) for k, v in keys.items()})


def make_dirichlet_fit():
    dd = load_by_path('ref_dirichletDifferentiation',
                      'xview/models/dirichletDifferentiation.py')
    ff = load_by_path('ref_dirichlet_fastfit', 'xview/models/dirichlet_fastfit.py')
    rng = np.random.default_rng(868)
    out = {}
    cases = []
    for i, (c, conc, delta, beta) in enumerate([(4, 2.0, 1e-2, 1e-2), (12, 0.7, 1e-2, 1e-2),
                                                (12, 5.0, 0.0, 0.0), (14, 1.5, 1e-3, 0.1),
                                                (10, 0.3, 1e-2, 0.5)]):
        alpha = rng.gamma(2.0, conc, size=c) + 0.2
        pos = rng.dirichlet(alpha, size=4000)
        neg = rng.dirichlet(np.ones(c), size=4000)
        ss = np.log(1e-10 + pos).mean(0)
        neg_ss = np.log(1e-10 + neg).mean(0)
        with contextlib.redirect_stdout(io.StringIO()):
            res = dd.findDirichletPriors(ss, neg_ss, np.ones(c).astype('float64'),
                                         max_iter=10000, delta=delta, beta=beta)
        out['fit%d_ss' % i] = ss
        out['fit%d_neg_ss' % i] = neg_ss
        out['fit%d_delta_beta' % i] = np.array([delta, beta])
        out['fit%d_alpha' % i] = np.asarray(res, np.float64)
        cases.append(i)
    out['fit_cases'] = np.array(cases)
    # fastfit pieces: moment init, inverse digamma, fixed point
    d = rng.dirichlet(np.array([3.0, 1.0, 0.5, 6.0, 2.0]), size=20)   # T=20 MC samples
    out['ff_D'] = d
    out['ff_init_a'] = ff._init_a(d)
    y = np.linspace(-6.0, 4.0, 41)
    out['ff_ipsi_y'] = y
    out['ff_ipsi_x'] = ff._ipsi(y)
    out['ff_fixedpoint'] = ff._fixedpoint(d, tol=1e-7, maxiter=1000)
    np.savez_compressed(os.path.join(HERE, 'dirichlet_fit.npz'), **out)
    print('dirichlet_fit.npz:', len(out), 'arrays')


if __name__ == '__main__':
    make_exp868()
    make_weight_keys()
    make_dirichlet_fit()
