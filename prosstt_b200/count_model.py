"""UMI count model: per-gene variance hyper-parameters and the NB parameterisation.

Mirror of the hot-path part of prosstt/count_model.py: generate_negbin_params (:14-48)
and get_pr_umi (:131-161).  The amplification-model pmfs of the reference
(lognegbin/negbin/my_negbin/sum_negbin/get_pr_amp, :51-228) are not on the simulation
path (their only use is commented out at simulation.py:649-650) and are out of scope.
"""
import numpy as np
import torch

from prosstt_b200 import _native as nat


def generate_negbin_params(tree, mean_alpha=0.2, mean_beta=2, a_scale=1.5, b_scale=1.5):
    """alpha_g = exp(N(log mean_alpha, log a_scale)), beta_g = exp(N(log mean_beta,
    log b_scale)) + 1; alpha draws first (count_model.py:42-48).  Host-side O(G) setup on
    the global legacy numpy stream, so np.random.seed(s) reproduces the reference's values."""
    G = tree.G
    alphas = np.exp(np.random.normal(loc=np.log(mean_alpha), scale=np.log(a_scale), size=G))
    betas = np.exp(np.random.normal(loc=np.log(mean_beta), scale=np.log(b_scale), size=G)) + 1
    return alphas, betas


def get_pr_umi(a, b, m, device=None):
    """(p, r) of the negative binomial with mean m and variance a*m^2 + b*m
    (count_model.py:156-161): p = (s2-m)/s2, r = m^2/(s2-m), both 0 where s2 <= 0.
    Evaluated on the GPU in fp64 (pst_nb_params); m may be (G,) or (N, G)."""
    dev = nat.device(device)
    m_arr = np.asarray(m, dtype=np.float64)
    G = m_arr.shape[-1] if m_arr.ndim else 1
    m2 = np.ascontiguousarray(m_arr.reshape(-1, G))
    a_g = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (G,)))
    b_g = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), (G,)))
    d_m = nat.to_dev(m2, torch.float64, dev)
    d_a = nat.to_dev(a_g, torch.float64, dev)
    d_b = nat.to_dev(b_g, torch.float64, dev)
    d_p = torch.empty_like(d_m)
    d_r = torch.empty_like(d_m)
    nat.call("pst_nb_params", nat.ptr(d_a), nat.ptr(d_b), nat.ptr(d_m), m2.shape[0], G,
             nat.ptr(d_p), nat.ptr(d_r), nat.stream_ptr(dev))
    return d_p.cpu().numpy().reshape(m_arr.shape), d_r.cpu().numpy().reshape(m_arr.shape)


def sampler_params_f32(a, b, m, scaling=1.0, device=None):
    """The fp32 parameterisation exactly as the count sampler evaluates it (pst_nb_params_f32 runs
    the __device__ functions that pst_draw_counts inlines, in the same order): for means-table
    values m, library sizes `scaling` and variance parameters a, b (broadcast to a common shape) a dict
    of mu = m s, theta = (a m) s + b - 1, r = mu/theta (scipy's n), q = theta/(1+theta) (get_pr_umi's p),
    a = q r, log2p0 = log2 P(X = 0) and route (0 mixture, 1 inversion, 2 domain error)."""
    dev = nat.device(device)
    m_arr, s_arr, a_arr, b_arr = np.broadcast_arrays(np.asarray(m, np.float64), np.asarray(scaling, np.float64),
                                                     np.asarray(a, np.float64), np.asarray(b, np.float64))
    shape = m_arr.shape
    n = int(m_arr.size)
    d_m = nat.to_dev(m_arr.ravel(), torch.float32, dev)
    d_s = nat.to_dev(s_arr.ravel(), torch.float32, dev)
    d_a = nat.to_dev(a_arr.ravel(), torch.float32, dev)
    d_b = nat.to_dev((b_arr - 1.0).ravel(), torch.float32, dev)       # beta-1 formed in fp64 like gene_params
    outs = {k: torch.empty(n, dtype=torch.float32, device=dev) for k in ("mu", "theta", "r", "q", "a", "log2p0")}
    route = torch.empty(n, dtype=torch.int32, device=dev)
    nat.call("pst_nb_params_f32", d_m, d_s, d_a, d_b, n, outs["mu"], outs["theta"], outs["r"], outs["q"], outs["a"],
             outs["log2p0"], route, nat.stream_ptr(dev))
    res = {k: v.cpu().numpy().reshape(shape) for k, v in outs.items()}
    res["route"] = route.cpu().numpy().reshape(shape)
    return res
