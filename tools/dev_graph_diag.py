"""Diagnostic: eager vs GraphedStep gradients per scene and per loss subset (c1 config)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import coocc_b200
from test_gpu_graph import _scene, _build, KEYS

coocc_b200.set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32")
subsets = {"all-lovasz": [k for k in KEYS if "lovasz" not in k], "ce": ["loss_voxel_ce_c_0"], "sem": ["loss_voxel_sem_scal_c_0"],
           "geo": ["loss_voxel_geo_scal_c_0"], "render": ["loss_depth_render", "loss_rgb"], "lovasz": ["loss_voxel_lovasz_c_0"]}
scenes = [_scene("c1", s) for s in (0, 1, 2, 0)]
for name, keys in subsets.items():
    m_e, _ = _build("c1"); m_2, _ = _build("c1"); m_g, _ = _build("c1")
    m_2.load_state_dict(copy.deepcopy(m_e.state_dict())); m_g.load_state_dict(copy.deepcopy(m_e.state_dict()))
    e1 = coocc_b200.GraphedStep(m_e, None, None, keys, enabled=False)
    e2 = coocc_b200.GraphedStep(m_2, None, None, keys, enabled=False)
    gr = coocc_b200.GraphedStep(m_g, None, None, keys, bucket=1 << 20)
    for i, sc in enumerate(scenes):
        a, b, c = float(e1(*sc)), float(e2(*sc)), float(gr(*sc))
        def err(ma, mb):
            num = den = 0.0; worst = []
            for (n, p), q in zip(ma.named_parameters(), mb.parameters()):
                if p.grad is None: continue
                d = float(((p.grad.double() - q.grad.double()) ** 2).sum()); num += d; den += float((p.grad.double() ** 2).sum())
                worst.append((d, n))
            worst.sort(reverse=True)
            return (num / max(den, 1e-300)) ** 0.5, [(n, "%.1e" % (d ** 0.5)) for d, n in worst[:3]], den ** 0.5
        en, _, _ = err(m_e, m_2)
        eg, w, nrm = err(m_e, m_g)
        print("%-10s scene %d  loss e=%.7f e2=%.7f g=%.7f | noise %.2e  graph %.2e  |g|=%.2e  worst %s" % (name, i, a, b, c, en, eg, nrm, w), flush=True)
