"""The oracle (oracle/prosstt_oracle.py) against fixtures produced by RUNNING the
reference (tests/golden/make_golden.py).  CPU only; pins the oracle before any
CUDA result is compared with it."""
import numpy as np
import pytest

from conftest import golden_lineage, load_maps, load_npz, pairs_to_dict
from oracle import prosstt_oracle as orc

MAPS = load_maps()


def _otree(rec):
    time = pairs_to_dict(rec["time"])
    return orc.OTree(rec["topology"], time)


@pytest.mark.parametrize("rec", MAPS, ids=[r["name"] for r in MAPS])
def test_integer_maps(rec):
    t = _otree(rec)
    assert t.branches == rec["branches"]
    assert t.root == rec["root"]
    bt = orc.branch_times(t)
    assert [[k, v] for k, v in bt.items()] == rec["branch_times"]
    zones = orc.populate_timezone(t)
    assert zones == rec["timezone"]
    assign = orc.assign_branches(bt, zones)
    assert [[k, v] for k, v in assign.items()] == rec["assignments"]
    pt, br = orc.cover_whole_tree(t)
    assert pt == rec["cover_pt"] and br == rec["cover_br"]
    assert orc.max_time(t) == rec["max_time"]
    assert [str(b) for b in orc.bfs_branches(t)] == [str(b) for b in rec["bfs"]]
    assert orc.paths(t, t.root) == rec["paths"]
    assert abs(sum(np.sum(v) for v in t.density.values()) - rec["density_sum"]) < 1e-15


def test_docstring_examples():
    # tree.py:386-390
    t = orc.OTree([[0, 1], [0, 2]], {0: 40, 1: 40, 2: 40})
    assert dict(orc.branch_times(t)) == {0: [0, 39], 1: [40, 79], 2: [40, 79]}
    # sim_utils.py:276-293
    t = orc.OTree([[0, 1], [0, 2], [2, 3], [2, 4]], {0: 10, 1: 20, 2: 8, 3: 25, 4: 15})
    a = orc.assign_branches(orc.branch_times(t), orc.populate_timezone(t))
    assert list(a.values()) == [[0], [1, 2], [1, 3, 4], [3, 4], [3]]
    # SURVEY.md 3.4 probe
    t = orc.OTree([["A", "B"], ["A", "C"]], {"A": 3, "B": 4, "C": 2})
    pt, br = orc.cover_whole_tree(t)
    assert pt == [0, 1, 2, 3, 4, 3, 4, 5, 6]
    assert br == list("AAABBCCBB")


@pytest.mark.parametrize("name", ["abc", "bp2", "fork"])
def test_lineage_from_draws(name):
    branches, time, top, d = golden_lineage(name)
    t = orc.OTree(top, time, G=int(d["G"]), modules=int(d["K"]))
    draws = {b: (d["u0_%s" % b], d["v0_%s" % b], d["eta_%s" % b], d["eps_%s" % b]) for b in branches}
    for b in branches:
        assert np.array_equal(orc.branch_programs_from_draws(*draws[b]), d["raw_%s" % b])
    W, rel = orc.lineage_from_draws(t, draws, d["H"])
    for b in branches:
        assert np.array_equal(W[b], d["W_%s" % b])
        assert np.array_equal(rel[b], d["rel_%s" % b])
        p = orc.parent_of(t, b)
        if p is not None:  # continuity at the fork
            assert np.allclose(W[b][0], W[p][-1], rtol=0, atol=1e-14)
    assert np.array_equal(orc.max_rel_exp(t, rel), d["max_rel_exp"])
    M = orc.absolute_means(rel, d["gene_scale"])
    for b in branches:
        assert np.array_equal(M[b], d["M_%s" % b])
    assert [str(b) for b in orc.bfs_branches(t)] == [str(b) for b in d["bfs"]]
    assert np.all(d["gene_scale"] * d["max_rel_exp"] <= 5000)


def test_base_gene_exp_replay():
    # sim_utils.py:463-469: redraw while scale*max > abs_max
    mx = np.array([1.0, 4000.0, 2.0])
    z = np.array([0.1, 3.0, 2.5, -1.0, 0.3])
    out, used = orc.base_gene_exp_from_normals(mx, z)
    assert used == 5 and np.allclose(out, np.exp([0.1, -1.0, 0.3]))


def test_nb_params():
    d = load_npz("nbparams.npz")
    p, r = orc.get_pr_umi(d["a"], d["b"], d["m"])
    assert np.array_equal(p, d["p"]) and np.array_equal(r, d["r"])
    p, r = orc.get_pr_umi(d["a2"], d["b2"], d["m"])
    assert np.array_equal(p, d["p2"]) and np.array_equal(r, d["r2"])
    a, b = orc.negbin_params_from_normals(d["gen_za"], d["gen_zb"])
    assert np.array_equal(a, d["gen_alpha"]) and np.array_equal(b, d["gen_beta"])
    # closed form used by the kernels: shape r = mu/theta,
    # scale theta = alpha*mu + beta - 1, scipy p = 1/(1+theta)
    theta = d["a"] * d["m"] + d["b"] - 1
    assert np.allclose(d["r"], d["m"] / theta, rtol=1e-9)
    assert np.allclose(1 - d["p"], 1 / (1 + theta), rtol=1e-9)
    # s2 <= 0 -> zeros (count_model.py:159-160)
    p0, r0 = orc.get_pr_umi([0.1], [2.0], [0.0])
    assert p0[0] == 0 and r0[0] == 0


def _sampling_tree(name):
    branches, time, top, d = golden_lineage(name)
    s = load_npz("sampling_%s.npz" % name)
    dens, o = {}, 0
    for b in branches:
        dens[b] = s["density"][o:o + time[b]]
        o += time[b]
    t = orc.OTree(top, time, G=int(d["G"]), modules=int(d["K"]), density=dens)
    t.means = {b: d["M_%s" % b] for b in branches}
    return t, s


@pytest.mark.parametrize("name", ["abc", "bp2", "fork"])
def test_sampler_index_maps_from_draws(name):
    t, s = _sampling_tree(name)
    pt, br, _ = orc.sample_density_index(t, s["dens_u"])
    assert np.array_equal(pt, s["dens_pt"])
    assert [str(b) for b in br] == [str(b) for b in s["dens_br"]]
    assert np.array_equal(orc.scalings_from_normals(s["dens_z"]), s["dens_scalings"])
    times = orc.draw_times_from_normals(s["ser_zt"], orc.max_time(t))
    assert np.array_equal(times, s["ser_pt"])
    br = orc.pick_branches_from_uniforms(t, times, s["ser_upick"])
    assert [str(b) for b in br] == [str(b) for b in s["ser_br"]]
    pt, br = orc.cover_whole_tree(t)
    assert np.array_equal(np.repeat(pt, int(s["wt_n"])), s["wt_pt"])
    assert [str(b) for b in np.repeat(np.array(br, dtype=object), int(s["wt_n"]))] == [str(b) for b in s["wt_br"]]


@pytest.mark.parametrize("name", ["abc", "bp2", "fork"])
def test_full_samplers_replay_legacy_stream(name):
    """Whole sampler runs, bit for bit, on the legacy MT19937 stream: pins
    draw_counts (gather, scaling, get_pr_umi, negative_binomial) end to end."""
    t, s = _sampling_tree(name)
    X, pt, br, sc = orc.sample_density(t, int(s["dens_N"]), s["alpha"], s["beta"],
                                       np.random.RandomState(int(s["dens_seed"])))
    assert np.array_equal(X, s["dens_X"]) and np.array_equal(sc, s["dens_scalings"])
    X, pt, br, sc = orc.sample_pseudotime_series(t, s["ser_cells"], s["ser_points"], s["ser_std"],
                                                 s["alpha"], s["beta"],
                                                 np.random.RandomState(int(s["ser_seed"])))
    assert np.array_equal(X, s["ser_X"]) and np.array_equal(pt, s["ser_pt"])
    X, pt, br, sc = orc.sample_whole_tree(t, int(s["wt_n"]), s["alpha"], s["beta"],
                                          np.random.RandomState(int(s["wt_seed"])))
    assert np.array_equal(X, s["wt_X"]) and np.array_equal(sc, s["wt_scalings"])


def test_timeseries_input_quirks():
    # sim_utils.py:529-537: int cells floor-divided per point, scalar
    # std divided by the number of points
    pts, cells, std = orc.timeseries_input([0, 10, 20], 100, 6.0)
    assert cells.tolist() == [33, 33, 33] and std.tolist() == [2.0, 2.0, 2.0]


def test_draw_counts_domain_error():
    t, s = _sampling_tree("abc")
    with pytest.raises(ValueError):
        orc.draw_counts(t, [0], [t.branches[0]], [0.0], s["alpha"], s["beta"], np.random.RandomState(0))


def test_pearson_matches_scipy():
    import scipy.stats
    rng = np.random.RandomState(3)
    a, b = rng.normal(size=(9, 30)), rng.normal(size=(12, 30))
    ref = sum(scipy.stats.pearsonr(a[:9, g], b[:9, g])[0] < 0 for g in range(30))
    assert orc.pearson_anticorrelated(a, b) == ref


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, out in kat:
        got = orc.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(x) for x in got) == out


@pytest.mark.parametrize("name", ["abc", "bp2", "fork"])
def test_oracle_simulate_lineage_replays_reference(name):
    """Whole simulate_lineage (incl. its rejection loop) on the legacy stream."""
    branches, time, top, d = golden_lineage(name)
    t = orc.OTree(top, time, G=int(d["G"]), modules=int(d["K"]))
    rel, W, H = orc.simulate_lineage(t, np.random.RandomState(int(d["seed"])), a=float(d["a"]),
                                     rel_exp_cutoff=float(d["cutoff"]))
    assert np.array_equal(H, d["H"])
    for b in branches:
        assert np.array_equal(W[b], d["W_%s" % b]) and np.array_equal(rel[b], d["rel_%s" % b])
