"""E-step at small K (narrow MMA tiles, tensor-memory operand variant) against the fp32 direct-form kernel."""
import sys
import torch
sys.path.insert(0, '.')
from scd_b200 import kmeans
bad_total = 0
for (n, d, k) in [(5000, 128, 1), (5000, 128, 2), (5000, 128, 12), (5000, 768, 16), (5000, 768, 17), (5000, 96, 33), (40000, 256, 5),
                  (300, 768, 3), (129, 104, 7)]:
    g = torch.Generator().manual_seed(n + d + k)
    X = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).cuda()
    C = X[torch.randperm(n, generator=g)[:k]].clone()
    ref = torch.empty(n, dtype=torch.int64, device='cuda'); lab = torch.empty_like(ref)
    md_ref = torch.empty(n, device='cuda'); md = torch.empty(n, device='cuda')
    acc = torch.zeros(1, dtype=torch.float64, device='cuda'); acc2 = torch.zeros(1, dtype=torch.float64, device='cuda')
    kmeans._estep(X, C, ref, acc, md_ref, exact=True)
    for _ in range(3):
        acc2.zero_()
        kmeans._estep(X, C, lab, acc2, md)
    torch.cuda.synchronize()
    bad = int((lab != ref).sum()); err = float((md - md_ref).abs().max())
    bad_total += bad > max(2, n // 2000) or err > 1e-4
    print(f'n={n} d={d} k={k}: label mismatches {bad}, max |mindist err| {err:.2e}, inertia rel err {abs(acc2.item() - acc.item()) / max(acc.item(), 1e-9):.2e}')
print('FAIL' if bad_total else 'all ok')
