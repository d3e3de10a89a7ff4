"""Per-parameter gradient differences: exact vs exact (run-to-run noise of the atomics) and graph replay vs exact."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import test_gpu_graph as T
from oracle.testing import rel_err
from virconv_b200.graph import GraphedStep

model = T._model('bf16')
batches = [T._batch([20 + 2 * i, 21 + 2 * i]) for i in range(3)]
ref = [T._exact(model, b)[:2] for b in batches]
ref2 = [T._exact(model, b)[:2] for b in batches]
step = GraphedStep(model, T._loss, margin=1.35, grain=256)
params = dict(model.named_parameters())
got = []
for s in range(6):
    loss = step(batches[s % 3])
    got.append((loss.detach().clone(), {k: v.grad.detach().clone() for k, v in params.items()}))
torch.cuda.synchronize()
print('recaptures', step.recaptures, 'err flag', T._err_flag())
for s in (0, 1, 2, 5):
    l0, g0 = ref[s % 3]
    print('step', s, 'loss', float(got[s][0]), l0)
    rows = []
    for k in params:
        e_noise = rel_err(ref2[s % 3][1][k].cpu(), g0[k].cpu())
        e_graph = rel_err(got[s][1][k].cpu(), g0[k].cpu())
        rows.append((e_graph, e_noise, k))
    rows.sort(reverse=True)
    for e_graph, e_noise, k in rows[:8]:
        print('   %-40s graph-vs-exact %.2e   exact-vs-exact %.2e' % (k, e_graph, e_noise))
