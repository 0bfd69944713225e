"""Repro helper: E-step launches at a given shape, checked against the fp32 direct-form kernel (exact=1)."""
import sys
import torch
sys.path.insert(0, '.')
from scd_b200 import kmeans

n, d, k, reps = (int(x) for x in sys.argv[1:5])
X = torch.randn(n, d, device='cuda'); X = X / X.norm(dim=1, keepdim=True)
C = X[:k].clone()
labels = torch.empty(n, dtype=torch.int64, device='cuda'); acc = torch.zeros(1, dtype=torch.float64, device='cuda')
ref = torch.empty_like(labels)
kmeans._estep(X, C, ref, acc, exact=True)
torch.cuda.synchronize()
for i in range(reps):
    kmeans._estep(X, C, labels, acc)
    torch.cuda.synchronize()
    print('launch', i, 'ok, label mismatches vs the fp32 direct form:', int((labels != ref).sum()), flush=True)
print('done')
