"""GPU: +I models (proportion of invariant sites, +IU{p}) through the C ABI - every kernel that
forms a site likelihood or a sumtable (tree log-likelihood, lookup build, lane = site BLO with and
without first-round tables, per-rate scalers, the 8-category DNA kernel, both amino-acid variants)
against the oracle, and the placements against the reference's recorded ones
(tests/golden/pinv, tests/golden/make_golden_pinv.py; the oracle's +I path is pinned on the same
vectors in tests/test_oracle_pinv.py)."""
import json
import os
import subprocess

import numpy as np
import pytest

import helpers
from test_oracle_pinv import CFG1_PINV, RATE300, RATE300_PINV, SYNTH64_PINV, SYNTHAA_PINV, gold

pytestmark = pytest.mark.gpu


def _file_case(name, model, tree="tree.nwk", ref="ref.fasta"):
    d = os.path.join(helpers.GOLDEN, name)
    return helpers.load_case(os.path.join(d, tree), os.path.join(d, ref), os.path.join(d, "query.fasta"), model)


def _check_stages(case, ctx, edge_step=1):
    """tree log-likelihood across edges and the lookup tables against the oracle"""
    want = case.ref.tree_logl(0)
    edges = range(0, case.tree.num_branches, edge_step)
    vals = [ctx.edge_loglikelihood(e) for e in edges]
    assert np.allclose(vals, want, rtol=1e-11, atol=0)
    lk = case.placer.lookup if case.placer.lookup is not None else case.placer.build_lookup()
    for e in edges:
        assert np.allclose(ctx.get_lookup(e), lk[e], rtol=1e-11, atol=1e-11), f"lookup of edge {e}"


def _check_pairs(case, ctx, opts, logl_rel=1e-8, len_abs=1e-5, every=1):
    ctx.upload_queries(case.query_rows)
    if opts.prescoring:
        ctx.preplace()
    ctx.select(opts)
    ctx.place_pairs(opts)
    q, e, raw = ctx.get_pairs()
    assert len(q) > 0
    for qi, ei, r in list(zip(q, e, raw))[::every]:
        p = case.placer.thorough(case.qseqs[qi], int(ei))
        assert abs(r["likelihood"] - p.logl) <= logl_rel * abs(p.logl), (qi, ei, r, p)
        assert abs(r["pendant_length"] - p.pendant) <= len_abs and abs(r["distal_length"] - p.distal) <= len_abs, (qi, ei, r, p)


def _check_placements(case, ctx, opts, want, logl_rel=1e-6):
    out, counts = ctx.place_chunk(case.query_rows, opts)
    got = dict(zip(case.qnames, helpers.records_to_lists(out, counts)))
    bad = []
    for name in want:
        try:
            helpers.assert_placements_close(got[name], want[name], name, logl_rel=logl_rel)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(want)} queries differ from the reference: {bad[:3]}"


@pytest.mark.parametrize("no_first", [False, True], ids=["first_tables", "full_first_pass"])
def test_cfg1_pinv(built, monkeypatch, no_first):
    if no_first:
        monkeypatch.setenv("EPA_B200_NO_FIRST", "1")       # read at context creation
    g = gold()
    case = helpers.cfg1_case(CFG1_PINV)
    assert (case.model.invariant >= 0).sum() > 100
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_stages(case, ctx)
    _check_pairs(case, ctx, built.capi.default_options(prescoring=0), logl_rel=1e-9, len_abs=1e-6)
    _check_placements(case, ctx, built.capi.default_options(), g["cfg1_default"]["placements"])
    _check_placements(case, ctx, built.capi.default_options(prescoring=0, support_threshold=0.0, filter_max=13),
                      g["cfg1_noheur_all"]["placements"])
    ctx.close()


def test_synth64_pinv(built):
    g = gold()
    case = _file_case("synth64", SYNTH64_PINV)
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_stages(case, ctx, edge_step=7)
    _check_pairs(case, ctx, built.capi.default_options(), every=5)
    _check_placements(case, ctx, built.capi.default_options(), g["synth64_default"]["placements"])
    ctx.close()


def test_synth64_pinv_eight_rate_categories(built):
    """+G8: the 4-lanes-per-site DNA kernel (kernels_blo.cuh) and the generic lookup build"""
    case = _file_case("synth64", "GTR{1/2/1/1/2/1}+FU{0.3/0.2/0.2/0.3}+IU{0.15}+G8{0.5}")
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_stages(case, ctx, edge_step=11)
    _check_pairs(case, ctx, built.capi.default_options(), every=9)
    ctx.close()


@pytest.mark.parametrize("per_rate", [False, True], ids=["site_scalers", "rate_scalers"])
def test_rate300_pinv(built, per_rate):
    """CLVs that do get rescaled, with per-site and with per-rate scalers (the reference's window offset)"""
    g = gold()
    ds = built.synth.dataset(**RATE300)
    case = helpers.case_from_arrays(ds["newick"], ds["names"], ds["ref"], ds["qnames"], ds["queries"], RATE300_PINV,
                                    per_rate=per_rate, bugcompat=per_rate, column_mask=True)
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_stages(case, ctx, edge_step=29)
    _check_pairs(case, ctx, built.capi.default_options())
    _check_placements(case, ctx, built.capi.default_options(),
                      g["rate300_rate" if per_rate else "rate300_site"]["placements"])
    ctx.close()


@pytest.mark.parametrize("old_aa", [False, True], ids=["unit_mapped", "site_rate_threads"])
def test_synthaa_pinv(built, monkeypatch, old_aa):
    if old_aa:
        monkeypatch.setenv("EPA_B200_OLD_AA", "1")
    g = gold()
    case = _file_case("synthaa", SYNTHAA_PINV)
    ctx = helpers.make_context(case)
    ctx.build_lookup()
    _check_stages(case, ctx, edge_step=5)
    _check_pairs(case, ctx, built.capi.default_options(), every=3)
    _check_placements(case, ctx, built.capi.default_options(), g["synthaa_default"]["placements"])
    ctx.close()


def test_host_layer_pinv_files_to_jplace(built, tmp_path):
    """model string with +IU{p} through the C++ host layer and the command-line program"""
    g = gold()
    d = os.path.join(helpers.GOLDEN, "cfg1")
    out = str(tmp_path / "lib")
    built.session.run_files(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), os.path.join(d, "query.fasta"),
                            CFG1_PINV, out)
    doc = json.load(open(os.path.join(out, "epa_result.jplace")))
    got = {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}
    for name, want in g["cfg1_default"]["placements"].items():
        helpers.assert_placements_close(got[name], want, name)
    exe = os.path.join(helpers.ROOT, "epa-ng_b200", "epa-ng-b200")
    d = os.path.join(helpers.GOLDEN, "synth64")
    out = str(tmp_path / "cli")
    subprocess.run([exe, "-t", os.path.join(d, "tree.nwk"), "-s", os.path.join(d, "ref.fasta"), "-q",
                    os.path.join(d, "query.fasta"), "-m", SYNTH64_PINV, "-w", out, "--redo"], check=True, stdout=subprocess.DEVNULL)
    doc = json.load(open(os.path.join(out, "epa_result.jplace")))
    got = {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}
    bad = []
    for name, want in g["synth64_default"]["placements"].items():
        try:
            helpers.assert_placements_close(got[name], want, name)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} differ: {bad[:3]}"
