#!/usr/bin/env python
"""Diagnostic: per-bin z-scores of the hybrid sampler against the exact pmf for one regime (N draws)."""
import argparse, sys
import numpy as np, scipy.stats, torch
sys.path.insert(0, ".")
from prosstt_b200 import tree as ptree
from prosstt_b200.device import CountEngine, TreeTables
ap = argparse.ArgumentParser()
ap.add_argument("--mu", type=float, default=20.0); ap.add_argument("--alpha", type=float, default=0.9)
ap.add_argument("--beta", type=float, default=2.0); ap.add_argument("--draws", type=float, default=1e9)
ap.add_argument("--sampler", default="hybrid"); ap.add_argument("--seed", type=int, default=991)
a = ap.parse_args()
dev = torch.device("cuda:0")
G = 128
t = ptree.Tree(topology=[["A", "B"]], time={"A": 1, "B": 1}, num_branches=2, branch_points=0, modules=1, G=G)
m = np.full((1, G), a.mu)
t.add_genes({"A": m.copy(), "B": m.copy()})
eng = CountEngine(t, TreeTables(t, dev), np.full(G, a.alpha), np.full(G, a.beta), dev, sampler=a.sampler)
N = int(a.draws); n = 4_000_000; done = 0
hist = torch.zeros(1, dtype=torch.int64, device=dev)
while done < N:
    X = eng.draw(torch.zeros(n, dtype=torch.int32, device=dev), torch.ones(n, dtype=torch.float32, device=dev), a.seed, done // G)
    h = torch.bincount(X.view(-1))
    if h.numel() > hist.numel():
        h[:hist.numel()] += hist; hist = h
    else:
        hist[:h.numel()] += h
    done += n * G
eng.check()
theta = a.alpha * a.mu + a.beta - 1
r, p = a.mu / theta, 1 / (1 + theta)
hist = hist.cpu().numpy().astype(float); N = done
ks = np.arange(len(hist)); pmf = scipy.stats.nbinom.pmf(ks, r, p); sf = scipy.stats.nbinom.sf(ks - 1, r, p)
z = (hist - pmf * N) / np.sqrt(np.maximum(pmf * N, 1e-9))
body = pmf >= 1e-6
print("N=%.2e chi2/dof body %.3f" % (N, (z[body] ** 2).sum() / body.sum()))
for k in ks[body]:
    flag = "  <<<" if abs(z[k]) > 3.5 else ""
    if abs(z[k]) > 2.5 or k % 20 == 0:
        print("k=%4d pmf=%.3e sf=%.3e obs/exp=%.5f z=%+.2f%s" % (k, pmf[k], sf[k], hist[k] / (pmf[k] * N), z[k], flag))
# cumulative view: where does the deviation accumulate?
cz = np.cumsum(hist - pmf * N)
for k in range(0, len(ks), max(1, len(ks) // 40)):
    print("cum k<=%4d excess %+9.0f (sd %.0f)" % (k, cz[k], np.sqrt(N * sf[k] * (1 - sf[k]) if sf[k] < 1 else 1)))
