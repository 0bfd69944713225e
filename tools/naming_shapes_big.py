"""Multi-item (several row blocks per CTA pair) scoring/top-k launches at various embedding widths, each in its own
process, checked against torch on a row sample."""
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == '--one':
    import torch
    sys.path.insert(0, '.')
    from scd_b200 import naming
    n, v, d, k = (int(x) for x in sys.argv[2:6])
    g = torch.Generator(device='cuda').manual_seed(n + v + d)
    X = torch.randn(n, d, device='cuda', generator=g); X = (X / X.norm(dim=1, keepdim=True)).bfloat16()
    W = torch.randn(v, d, device='cuda', generator=g); W = (W / W.norm(dim=1, keepdim=True)).bfloat16()
    vocab = naming.Vocabulary.from_rows(W)
    vals, idx, _, _ = naming.name_topk_raw(X, vocab, k, False)
    torch.cuda.synchronize()
    rows = torch.randperm(n, device='cuda')[:2048]
    rv, ri = (100. * (X[rows].float() @ W.float().t())).topk(min(k, v), 1, True, True)
    print(f'n={n} v={v} d={d} k={k}: idx agree {float((idx[rows][:, :ri.shape[1]] == ri).float().mean()):.5f} '
          f'max val err {float((vals[rows][:, :ri.shape[1]] - rv).abs().max()):.2e}')
else:
    for (n, v, d, k) in [(40000, 11000, 64, 5), (40000, 11000, 128, 5), (40000, 11000, 192, 5), (40000, 11000, 512, 5),
                         (40000, 11000, 704, 5), (40000, 11000, 72, 5), (40000, 300, 64, 1), (127000, 100, 512, 1), (127000, 100, 768, 1), (60000, 224, 64, 5),
                         (60000, 500, 40, 5), (127000, 21000, 64, 5), (300000, 120, 768, 1)]:
        r = subprocess.run([sys.executable, __file__, '--one', str(n), str(v), str(d), str(k)], capture_output=True, text=True, timeout=120)
        out = (r.stdout + r.stderr).strip().splitlines()
        if r.returncode == 0:
            print(out[-1], flush=True)
        else:
            tags = sorted({l.split('tag')[1].split()[0] for l in out if 'timed out' in l})
            print(f'n={n} v={v} d={d} k={k}: FAILED rc={r.returncode}; timed-out wait tags {tags}; first: {[l for l in out if "timed out" in l][:2]}', flush=True)
