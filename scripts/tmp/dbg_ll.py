import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests'))
import torch
from util import make_model
from test_gpu_scene_ll import _inputs, _run
cases = [(dict(num_obj=9, width=50, height=50, max_obj_scale=0.22, debug_match_objects='greedy'), 1),
         (dict(num_obj=9, width=50, height=50, max_obj_scale=0.22, debug_match_objects='greedy'), 300),
         (dict(num_obj=9, width=50, height=50, max_obj_scale=0.22, debug_match_objects='greedy'), 1792),
         (dict(num_obj=6, width=50, height=50, max_obj_scale=0.22, debug_match_objects='greedy'), 1792),
         ({}, 4000), (dict(num_obj=9, debug_match_objects='greedy'), 1792)]
for kw, F_ in cases:
    oc, sd, model = make_model(kw, 31)
    img, z, w = _inputs(oc, F_, 5)
    a = _run(model, img, z, w, True)
    b = _run(model, img, z, w, False)
    d = (a['gz'] - b['gz']).abs()
    sc = b['gz'].abs().max()
    bad = (d > 2e-4 * sc).nonzero()
    print(kw.get('num_obj', 3), kw.get('width', 32), F_, 'gz rel err %.3e' % float(d.max() / sc), 'bad entries', len(bad))
    if len(bad):
        fr = bad[:, 0].unique()
        print('   bad frames (first 20):', fr[:20].tolist(), 'objects:', bad[:, 1].unique().tolist(), 'comps:', bad[:, 2].unique().tolist())
        f = int(fr[0])
        print('   z', z[f].tolist())
        print('   fused', a['gz'][f].tolist())
        print('   unfused', b['gz'][f].tolist())
