"""CPU checks of the arithmetic the count kernels are built from (no GPU): the constant tables are
read out of the shipped CUDA source and compared with their definitions, and the closed forms the
kernels use (series, Stirling bound, the t_k = P(k) k! inversion, the Marsaglia-Tsang acceptance
series, the NB parameterisation of count_model.py:156-161 in its gamma-Poisson form) are restated in
NumPy and compared with SciPy.  The GPU parity tests check the results of the kernels; these pin the
numbers and identities they rest on."""
import math
import os
import re

import numpy as np
import scipy.special
import scipy.stats

from conftest import ROOT

CSRC = os.path.join(ROOT, "prosstt_b200", "csrc")


def _source(name):
    with open(os.path.join(CSRC, name)) as fh:
        return fh.read()


def _floats(text):
    return [float(x) for x in re.findall(r"(-?\d+\.\d*(?:e[+-]?\d+)?)f", text)]


def test_factorial_tables_in_the_cuda_source():
    src = _source("pst_counts.cu")
    table = src[src.index("#define PST_INV_FACT_TABLE"):]
    table = table[:table.index("}") + 1].replace("\\\n", " ")
    inv_fact = _floats(table)
    assert len(inv_fact) == 34
    for k, v in enumerate(inv_fact):
        assert abs(v * math.factorial(k) - 1.0) < 2e-9, (k, v)
    logfact = src[src.index("c_logfact[16] = {"):]
    logfact = _floats(logfact[:logfact.index("}")])
    assert len(logfact) == 16
    for k, v in enumerate(logfact):
        assert abs(v - math.lgamma(k + 1)) < 1e-9, (k, v)
    # the head's unrolled terms stay inside the fp32 range: t_k = P(k) k! <= k! with mean <= 32 needs k <= 33
    assert math.factorial(33) < 3.4e38 / 32 and float(np.float32(inv_fact[33])) > 1.17549435e-38


def test_inversion_in_the_t_form_counts_like_the_quantile_function():
    """X = #{k : u > cdf(k)} with cdf built from t_k = P(k) k! (t_{k+1} = t_k (a + q k), NB; t_{k+1} = t_k lambda,
    Poisson) and the 1/k! table: equal to SciPy's ppf wherever u is not within rounding of a cdf step."""
    rng = np.random.RandomState(1)
    ks = np.arange(0, 34)
    inv_fact = np.array([1.0 / math.factorial(int(k)) for k in ks])
    for _ in range(200):
        mu, alpha, beta = np.exp(rng.uniform(-3, 2.7)), np.exp(rng.uniform(-4, 0)), 1 + np.exp(rng.uniform(-2, 1))
        theta = alpha * mu + beta - 1                       # count_model.py:156-161 in gamma-Poisson form
        r, q = mu / theta, theta / (1 + theta)
        a = q * r
        t = np.empty(34)
        t[0] = (1 + theta) ** -r
        for k in range(33):
            t[k + 1] = t[k] * (a + q * k)
        pmf = t * inv_fact
        assert np.allclose(pmf, scipy.stats.nbinom.pmf(ks, r, 1 / (1 + theta)), rtol=1e-10, atol=1e-300)
        cdf = np.cumsum(pmf)
        u = rng.uniform(0, min(cdf[-1], 1.0) * 0.999999, size=50)
        u = u[np.min(np.abs(u[:, None] - cdf[None, :]), axis=1) > 1e-9]
        assert np.array_equal((u[:, None] > cdf[None, :]).sum(axis=1), scipy.stats.nbinom.ppf(u, r, 1 / (1 + theta)))
    for lam in (1e-3, 0.3, 2.0, 9.99):                       # the gamma_poisson kernel's small-lambda stage
        t = np.exp(-lam) * lam ** ks
        cdf = np.cumsum(t * inv_fact)
        u = rng.uniform(0, 0.999, size=200)
        u = u[np.min(np.abs(u[:, None] - cdf[None, :]), axis=1) > 1e-9]
        assert np.array_equal((u[:, None] > cdf[None, :]).sum(axis=1), scipy.stats.poisson.ppf(u, lam))


def test_series_the_kernels_use_instead_of_cancelling_differences():
    # Poisson-limit log2 P(0) = -mu log2(e) log1p(theta)/theta (pst_common.cuh: nb_log2p0_small_theta, theta < 0.1)
    src = _source("pst_common.cuh")
    body = src[src.index("nb_log2p0_small_theta(float mu, float th) {"):]
    coef = _floats(body[:body.index("return")])
    want = [1.0 / 7, -1.0 / 6, 0.2, -0.25, 1.0 / 3, -0.5, 1.0]          # innermost first
    assert np.allclose(coef, want, atol=1e-9), coef
    th = np.linspace(1e-6, 0.1, 50)
    ser = np.polyval(coef, th)
    assert np.max(np.abs(ser / (np.log1p(th) / th) - 1)) < 2e-8          # truncation error at theta = 0.1
    # Marsaglia-Tsang acceptance bound 0.5 x^2 + d (1 - v + log v), v = (1 + e)^3, x = e sqrt(9 d): for |e| < 0.1
    # the kernels use d e^4 (-3/4 + 3/5 e - 1/2 e^2 + 3/7 e^3 - 3/8 e^4 + 1/3 e^5)
    cnt = _source("pst_counts.cu")
    line = cnt[cnt.index("h = d * e2 * e2 * ("):]
    poly = _floats(line[:line.index(";")])
    assert np.allclose(poly, [-0.75, 0.6, -0.5, 3.0 / 7, -0.375, 1.0 / 3], atol=1e-9), poly
    for d in (0.7, 5.0, 90.0):
        e = np.linspace(-0.1, 0.1, 41)
        e = e[e != 0]
        v = (1 + e) ** 3
        exact = 0.5 * 9 * d * e ** 2 + d * (1 - v + np.log(v))
        series = d * e ** 4 * np.polyval(poly[::-1], e)
        assert np.max(np.abs(series - exact)) < 4e-11 * d + 1e-13, d
    # PTRS acceptance bound -lam + k log(lam) - log(k!) as k (log1p(y) - y) - log sqrt(2 pi k) - 1/(12k) + 1/(360k^3)
    for lam in (10.0, 37.5, 400.0, 1e5):
        for k in (16.0, round(lam), round(lam + 4 * math.sqrt(lam)), round(max(16.0, lam - 4 * math.sqrt(lam)))):
            y = (lam - k) / k
            stirling = k * (math.log1p(y) - y) - 0.5 * math.log(2 * math.pi * k) - (1 / 12.0 - 1 / 360.0 / k ** 2) / k
            exact = -lam + k * math.log(lam) - math.lgamma(k + 1)
            assert abs(stirling - exact) < 1e-7 * max(1.0, abs(exact)), (lam, k)
    l1 = cnt[cnt.index("l1 = y2 * ("):]
    c = _floats(l1[:l1.index(";")])
    assert np.allclose(c, [(-1) ** (i + 1) / (i + 2) for i in range(len(c))], atol=1e-9), c   # log1p(y) - y
    y = np.linspace(-0.25, 0.25, 51)
    # truncated after y^11: 6.5e-9 at y = -1/4 (1.7e-7 of the value, fp32 rounding level), far less inside
    assert np.max(np.abs(y ** 2 * np.polyval(c[::-1], y) - (np.log1p(y) - y))) < 1e-8


def test_ptrs_trials_are_numbered_the_same_whether_taken_one_or_two_per_visit():
    """draw_counts_mixture_kernel visits a PTRS entry once per trial: trial t reads Philox block (t | 1) of the
    count's stream, words (x, y) when t is even and (z, w) when it is odd - the blocks 2a + 1 and word pairs of
    'two trials per block a'."""
    for a in range(64):
        for i in (0, 1):
            t = 2 * a + i
            assert (t | 1) == 2 * a + 1 and (t & 1) == i


def _ptrs_constants(src):
    """Every place the PTRS set-up is written in pst_counts.cu (mixture_step and the pipeline's PTRS stage)."""
    found = []
    for m in re.finditer(r"const float b = fmaf\(([\d.]+)f, slam, ([\d.]+)f\);\s*"
                         r"const float a = fmaf\(([\d.]+)f, b, (-?[\d.]+)f\);\s*"
                         r"const float log_inv_alpha = log_fast\(([\d.]+)f \+ div_fast\(([\d.]+)f, b - ([\d.]+)f\)\);\s*"
                         r"const float vr = ([\d.]+)f - div_fast\(([\d.]+)f, b - ([\d.]+)f\);", src):
        found.append(tuple(float(x) for x in m.groups()))
    return found


def test_ptrs_and_marsaglia_tsang_constants_sample_the_right_distributions():
    """Hoermann's PTRS (numpy/random/src/legacy + distributions.c: random_poisson_ptrs) and Marsaglia-Tsang
    restated in NumPy WITH THE CONSTANTS READ FROM THE CUDA SOURCE: chi-square against the Poisson pmf and KS
    against the gamma cdf.  A mistyped constant in either kernel copy fails here without a GPU."""
    src = _source("pst_counts.cu")
    sets = _ptrs_constants(src)
    assert len(sets) == 2 and sets[0] == sets[1], sets                     # both copies, same numbers
    c_b1, c_b0, c_a1, c_a0, c_ia0, c_ia1, c_ia2, c_vr0, c_vr1, c_vr2 = sets[0]
    trial = src[src.index("__device__ __forceinline__ bool ptrs_trial("):]
    trial = trial[:trial.index("float bound;")]
    shift, = re.findall(r"lam \+ ([\d.]+)f\)\)", trial)
    us_hi, = re.findall(r"us >= ([\d.]+)f && V <= vr", trial)
    us_lo, = re.findall(r"us < ([\d.]+)f && V > us", trial)
    shift, us_hi, us_lo = float(shift), float(us_hi), float(us_lo)
    rng = np.random.RandomState(7)
    for lam in (10.0, 33.0, 250.0):
        slam, loglam = math.sqrt(lam), math.log(lam)
        b = c_b0 + c_b1 * slam
        a = c_a0 + c_a1 * b
        inv_alpha = c_ia0 + c_ia1 / (b - c_ia2)
        vr = c_vr0 - c_vr1 / (b - c_vr2)
        n = 200000
        out = np.full(n, -1.0)
        todo = np.arange(n)
        while todo.size:
            U = rng.random_sample(todo.size) - 0.5
            V = rng.random_sample(todo.size)
            us = 0.5 - np.abs(U)
            k = np.floor((2 * a / us + b) * U + lam + shift)
            quick = (us >= us_hi) & (V <= vr)
            dead = (k < 0) | ((us < us_lo) & (V > us))
            with np.errstate(invalid="ignore", divide="ignore"):
                full = (np.log(V) + math.log(inv_alpha) - np.log(a / (us * us) + b)
                        <= -lam + k * loglam - scipy.special.gammaln(k + 1))
            ok = quick | (~dead & full)
            out[todo[ok]] = k[ok]
            todo = todo[~ok]
        lo, hi = int(scipy.stats.poisson.ppf(1e-4, lam)), int(scipy.stats.poisson.ppf(1 - 1e-4, lam))
        edges = np.arange(lo, hi + 2) - 0.5
        obs = np.histogram(np.clip(out, lo, hi), bins=edges)[0]
        pmf = scipy.stats.poisson.pmf(np.arange(lo, hi + 1), lam)
        pmf[0] += scipy.stats.poisson.cdf(lo - 1, lam)
        pmf[-1] += scipy.stats.poisson.sf(hi, lam)
        chi2 = ((obs - n * pmf) ** 2 / (n * pmf)).sum()
        assert scipy.stats.chi2.sf(chi2, len(pmf) - 1) > 1e-4, (lam, chi2)
    # Marsaglia-Tsang: d = shape - 1/3, c = 1/sqrt(9 d), v = (1 + c x)^3, squeeze u < 1 - 0.0331 x^4
    squeeze = {float(x) for x in re.findall(r"1\.0f - ([\d.]+)f \* x2 \* x2", src)}
    assert squeeze == {0.0331}
    for shape in (1.0, 2.5, 90.0):
        d = shape - 1.0 / 3
        c = 1 / math.sqrt(9 * d)
        x = rng.standard_normal(400000)
        u = rng.random_sample(x.size)
        v = (1 + c * x) ** 3
        with np.errstate(invalid="ignore", divide="ignore"):
            ok = (v > 0) & ((u < 1 - 0.0331 * x ** 4) | (np.log(u) < 0.5 * x * x + d * (1 - v + np.log(v))))
        g = d * v[ok]
        assert ok.mean() > 0.9
        assert scipy.stats.kstest(g[:100000], scipy.stats.gamma(shape).cdf).pvalue > 1e-4, shape


def test_pipeline_queues_never_overflow_and_finish_every_count():
    """Model of draw_counts_mixture_kernel's queue discipline (capacity read from the CUDA source): per cell
    iteration a warp appends up to 4 x 32 routed entries, then `run_queues(32)` drains full batches - small-lambda
    and PTRS first, then ONE batch of gamma retries, repeated - and `run_queues(1)` empties everything at the end.
    Whatever the mix of outcomes (all small, all PTRS, all rejected, random), no queue ever holds more than its
    capacity and every count is written exactly once."""
    src = _source("pst_counts.cu")
    cap = int(re.search(r"constexpr int MX_CAP = (\d+);", src).group(1))
    rng = np.random.RandomState(11)

    def simulate(p_reject, p_small, p_ptrs_reject, iterations):
        fill = {"g": [], "s": [], "l": []}                   # entries: (count id, attempt)
        high = {"g": 0, "s": 0, "l": 0}
        written = {}
        next_id = [0]

        def note():
            for k in fill:
                high[k] = max(high[k], len(fill[k]))
                assert len(fill[k]) <= cap, (k, len(fill[k]))

        def route(entries):                                  # outcome of one gamma attempt per entry
            for cid, att in entries:
                if att < 62 and rng.random_sample() < p_reject:
                    fill["g"].append((cid, att + 1))
                elif rng.random_sample() < p_small:
                    fill["s"].append((cid, 0))
                else:
                    fill["l"].append((cid, 0))
            note()

        def take(k, n):
            batch = fill[k][len(fill[k]) - n:]
            del fill[k][len(fill[k]) - n:]
            return batch

        def run_queues(limit):
            while any(len(fill[k]) >= limit for k in fill):
                while len(fill["s"]) >= limit:
                    for cid, _ in take("s", min(32, len(fill["s"]))):
                        written[cid] = written.get(cid, 0) + 1
                while len(fill["l"]) >= limit:
                    for cid, att in take("l", min(32, len(fill["l"]))):
                        if att < 125 and rng.random_sample() < p_ptrs_reject:
                            fill["l"].append((cid, att + 1))
                        else:
                            written[cid] = written.get(cid, 0) + 1
                    note()
                if len(fill["g"]) >= limit:
                    route(take("g", min(32, len(fill["g"]))))

        for _ in range(iterations):
            for j in range(4):                               # the four counts of a lane's quad, 32 lanes each
                ids = list(range(next_id[0], next_id[0] + 32))
                next_id[0] += 32
                route([(c, 0) for c in ids])
            run_queues(32)
            assert all(len(fill[k]) < 32 for k in fill)      # what the next iteration's 128 appends rely on
        run_queues(1)
        assert all(len(fill[k]) == 0 for k in fill)
        assert len(written) == next_id[0] and set(written.values()) == {1}
        return high

    assert 31 + 128 < cap + 1                                # 31 carried + 4 x 32 new entries fit
    for p_reject, p_small, p_ptrs in ((0.0, 1.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.9), (0.95, 0.5, 0.5),
                                      (0.05, 0.85, 0.12), (0.5, 0.0, 0.5), (1.0, 0.5, 0.0)):
        high = simulate(p_reject, p_small, p_ptrs, 40)
        assert max(high.values()) <= 31 + 128, high
