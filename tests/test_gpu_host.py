"""GPU: the C++ host layer end to end - files in, jplace out - against the reference's recorded
placements, through both the library entry point and the command-line program."""
import json
import os
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _read_jplace(path):
    doc = json.load(open(path))
    return {n: pq["p"] for pq in doc["placements"] for n in pq["n"]}, doc


def _check(got, want, what):
    assert set(got) == set(want)
    bad = []
    for name in want:
        try:
            helpers.assert_placements_close(got[name], want[name], f"{what}/{name}")
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, f"{len(bad)} of {len(want)} differ: {bad[:3]}"


def test_run_files_cfg1_matches_reference(built, tmp_path):
    d = os.path.join(helpers.GOLDEN, "cfg1")
    gold = helpers.golden("cfg1")
    for mname, model in (("gtrg", helpers.GTRG), ("gtrb", helpers.GTR_B)):
        out = str(tmp_path / mname)
        built.session.run_files(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), os.path.join(d, "query.fasta"),
                                model, out)
        got, doc = _read_jplace(os.path.join(out, "epa_result.jplace"))
        _check(got, gold[f"{mname}_default"]["placements"], mname)
        assert doc["tree"] == gold[f"{mname}_default"]["tree"]
        assert doc["version"] == 3
        assert doc["fields"] == ["edge_num", "likelihood", "like_weight_ratio", "distal_length", "pendant_length"]
        assert os.path.exists(os.path.join(out, "epa_info.log"))


def test_cli_synth64_matches_reference(built, tmp_path):
    d = os.path.join(helpers.GOLDEN, "synth64")
    exe = os.path.join(helpers.ROOT, "epa-ng_b200", "epa-ng-b200")
    assert os.path.exists(exe), "host program not built"
    model = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}"
    out = str(tmp_path / "cli")
    subprocess.run([exe, "-t", os.path.join(d, "tree.nwk"), "-s", os.path.join(d, "ref.fasta"), "-q",
                    os.path.join(d, "query.fasta"), "-m", model, "-w", out, "-T", "4", "--redo", "--chunk-size", "64"],
                   check=True, stdout=subprocess.DEVNULL)
    got, doc = _read_jplace(os.path.join(out, "epa_result.jplace"))
    gold = helpers.golden("synth64")["default"]
    _check(got, gold["placements"], "cli")
    assert doc["tree"] == gold["tree"]
    # the query order of the file is kept
    names = [pq["n"][0] for pq in doc["placements"]]
    assert names == sorted(names)
    # unsupported modes are refused by name, not silently ignored
    r = subprocess.run([exe, "-t", "x", "-s", "x", "-q", "x", "--dump-binary"], capture_output=True, text=True)
    assert r.returncode != 0 and "not supported" in r.stderr


def test_session_chunking_is_invisible(built):
    case = helpers.synth64_case()
    d = os.path.join(helpers.GOLDEN, "synth64")
    o = helpers.oracle()
    rn, rs = o.read_fasta(os.path.join(d, "ref.fasta"))
    _, qs = o.read_fasta(os.path.join(d, "query.fasta"))
    rs = o.apply_mask(rs, o.gap_mask(rs) | o.gap_mask(qs))        # the same pre-masking the case applied
    ref_rows = np.frombuffer("".join(rs).encode(), dtype=np.uint8).reshape(len(rs), -1)
    sess = built.session.Session(open(os.path.join(d, "tree.nwk")).read(), rn, ref_rows,
                                 "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}")
    assert abs(sess.tree_logl() - case.ref.tree_logl(0)) <= 1e-10 * abs(case.ref.tree_logl(0))
    a, ca = sess.place(case.query_rows, chunk_size=0)
    b, cb = sess.place(case.query_rows, chunk_size=37)
    assert np.array_equal(ca, cb) and np.array_equal(a, b)       # bit-identical whatever the chunking
    opts = built.capi.default_options(prescoring=0, support_threshold=0.0, filter_max=5)
    c, cc = sess.place(case.query_rows[:7], opts)
    gold = helpers.golden("synth64")["default"]["placements"]
    for qi in range(7):
        best = helpers.records_to_lists(c, cc)[qi][0]
        want = gold[case.qnames[qi]][0]
        assert best[0] == want[0] and abs(best[1] - want[1]) <= 1e-6 * abs(want[1])
    sess.close()


@pytest.mark.parametrize("fname", ["ref_rooted.tre", "ref_rooted_2.tre", "ref_rooted_3.tre", "ref_rooted_innerlabels.tre"])
def test_rooted_trees_match_reference(built, tmp_path, fname):
    """Rooted reference trees: unrooted for the computation, reported on the rooted tree
    (file_io.cpp:129-173, rtree_mapper.hpp:38-61); the MSA holds two taxa the tree does not."""
    d = os.path.join(helpers.GOLDEN, "cfg1")
    gold = json.load(open(os.path.join(d, "reference_rooted.json")))[fname]
    capi = built.capi
    runs = {"default": (capi.default_options(), True),
            "noheur_all": (capi.default_options(prescoring=0, support_threshold=0.0, filter_max=10), True),
            "unrooted": (capi.default_options(), False)}
    for rname, (opts, preserve) in runs.items():
        out = str(tmp_path / rname)
        built.session.run_files(os.path.join(d, fname), os.path.join(d, "aln.fasta"), os.path.join(d, "query.fasta"),
                                helpers.GTRG, out, opts=opts, preserve_rooting=preserve)
        got, doc = _read_jplace(os.path.join(out, "epa_result.jplace"))
        assert doc["tree"] == gold[rname]["tree"], rname
        _check(got, gold[rname]["placements"], f"{fname}/{rname}")


def test_bfast_query_file_gives_the_same_jplace(built, tmp_path):
    """A query file in the reference's binary 4-bit format (written by `epa-ng --bfast`) places like the FASTA."""
    d = os.path.join(helpers.GOLDEN, "cfg1")
    outs = {}
    for kind, q in (("fasta", "query.fasta"), ("bfast", "query.fasta.bfast")):
        out = str(tmp_path / kind)
        built.session.run_files(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), os.path.join(d, q), helpers.GTRG, out)
        outs[kind], _ = _read_jplace(os.path.join(out, "epa_result.jplace"))
    assert outs["fasta"] == outs["bfast"]
    _check(outs["bfast"], helpers.golden("cfg1")["gtrg_default"]["placements"], "bfast")


def test_run_files_empirical_frequencies(built, tmp_path):
    """+FC model string through the host layer: frequencies counted on the reference MSA, placements vs the
    reference's recorded ones (tests/golden/make_golden_freqs.py)."""
    d = os.path.join(helpers.GOLDEN, "cfg1")
    g = json.load(open(os.path.join(d, "reference_empirical.json")))
    for key in ("cfg1_fc_default", "cfg1_f_ic_default"):
        out = str(tmp_path / key)
        built.session.run_files(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), os.path.join(d, "query.fasta"),
                                g[key]["model"], out)
        got, _ = _read_jplace(os.path.join(out, "epa_result.jplace"))
        _check(got, g[key]["placements"], key)


def test_pipeline_many_chunks_and_threads_give_the_same_file(built, tmp_path):
    """epa_run_files_multi: the jplace does not depend on the chunk size, the number of host threads or
    the number of GPUs (one host thread per device, chunks handed out in file order)."""
    d = os.path.join(helpers.GOLDEN, "synth64")
    model = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}"
    files = (os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, "query.fasta"))
    texts = {}
    runs = {"one_chunk": dict(chunk_size=0, host_threads=1), "chunks_of_7": dict(chunk_size=7, host_threads=5),
            "chunks_of_64": dict(chunk_size=64, host_threads=0)}
    import torch
    if torch.cuda.device_count() >= 2:
        runs["two_gpus"] = dict(chunk_size=16, host_threads=4, devices=(0, 1))
    for name, kw in runs.items():
        out = str(tmp_path / name)
        st = built.session.run_files_multi(*files, model, out, invocation="pipeline test", **kw)
        assert st["n_queries"] == 200 and st["seconds_total"] > 0
        texts[name] = open(os.path.join(out, "epa_result.jplace")).read()
    first = texts["one_chunk"]
    for name, t in texts.items():
        assert t == first, name
    got, doc = _read_jplace(os.path.join(str(tmp_path / "chunks_of_7"), "epa_result.jplace"))
    _check(got, helpers.golden("synth64")["default"]["placements"], "pipeline")
    assert [pq["n"][0] for pq in doc["placements"]] == sorted(got)          # input order kept
    built.session.lib().epa_host_release_pinned_pool()


def test_pipeline_reports_query_errors(built, tmp_path):
    d = os.path.join(helpers.GOLDEN, "cfg1")
    bad = str(tmp_path / "bad.fasta")
    rows = open(os.path.join(d, "query.fasta")).read().replace("A", "!", 1)
    open(bad, "w").write(rows)
    with pytest.raises(built.capi.EpaError):
        built.session.run_files_multi(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), bad, helpers.GTRG, str(tmp_path / "o"))
    with pytest.raises(built.capi.EpaError, match="different widths|equal size"):
        short = str(tmp_path / "short.fasta")
        open(short, "w").write(">x\nACGT\n")
        built.session.run_files_multi(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), short, helpers.GTRG, str(tmp_path / "o2"))


def test_device_block_cache_reuse_and_trim(built):
    """Contexts allocate through the per-process device block cache: a second context of the same shape takes the
    blocks of the first, epa_device_pool_trim hands them back to the driver, and results do not depend on either."""
    import torch
    case = helpers.synth64_case()
    opts = built.capi.default_options()

    def run():
        ctx = helpers.make_context(case)
        ctx.build_lookup()
        out, counts = ctx.place_chunk(case.query_rows, opts)
        ctx.close()
        return out.copy(), counts.copy()

    a, ca = run()
    free_cached, _ = torch.cuda.mem_get_info(0)
    b, cb = run()                                   # reuses the cached blocks
    built.capi.load().epa_device_pool_trim()
    free_trimmed, _ = torch.cuda.mem_get_info(0)
    c, cc = run()                                   # allocates afresh
    assert np.array_equal(ca, cb) and np.array_equal(ca, cc)
    assert a.tobytes() == b.tobytes() == c.tobytes()
    assert free_trimmed >= free_cached              # the cache gave its blocks back
    built.capi.load().epa_device_pool_trim()
