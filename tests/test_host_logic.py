"""CPU: the C++ host layer (tree numbering, pruning schedule, model parsing) against the oracle and
the reference's recorded strings. No device is touched."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers


@pytest.fixture(scope="module")
def sess(built):
    return built.session


def _read(*p):
    return open(os.path.join(helpers.GOLDEN, *p)).read()


def test_numbered_newick_matches_reference(sess):
    # golden strings come from the unmodified reference (jplace "tree" field)
    nwk, nt, ne = sess.parse_tree(_read("cfg1", "ref.tre"))
    assert (nt, ne) == (8, 13)
    assert nwk == helpers.golden("cfg1")["gtrg_default"]["tree"]
    nwk, nt, ne = sess.parse_tree(_read("synth64", "tree.nwk"))
    assert (nt, ne) == (64, 125)
    assert nwk == helpers.golden("synth64")["default"]["tree"]


def test_tree_errors(sess, built):
    for bad in ["(A:1,B:1);", "(A:1,B:1,(C:1,D:1,E:1):1);", "(A:1,B:1,C:1", "A;", "((A:1,B:1,C:1):1,D:1);"]:
        with pytest.raises(built.capi.EpaError):
            sess.parse_tree(bad)
    # missing / zero lengths fall back to -ln(0.9) (set_missing_branch_lengths)
    nwk, _, _ = sess.parse_tree("(A,B:0,(C:0.5,D:0.25)x:0.125);", precision=4)
    assert nwk == "(A:0.1054{0},B:0.1054{1},(C:0.5000{2},D:0.2500{3})x:0.1250{4});"


@pytest.mark.parametrize("model", [helpers.GTRG, helpers.GTR_B, "GTR+G", "JC", "HKY{1.0/2.5}+FU{0.1/0.2/0.3/0.4}+G8{0.3}",
                                   "GTR{1/2/3/4/5/6}+FE+G4a{2.0}", "K80{1.0/4.0}"])
def test_model_matches_oracle(sess, model):
    o = helpers.oracle()
    want = o.parse_model(model)
    got = sess.parse_model(model)
    assert got["states"] == want.states and got["rate_cats"] == want.rate_cats
    assert np.allclose(got["rates"], want.rates, rtol=1e-13, atol=0)
    assert np.allclose(got["freqs"], want.freqs, rtol=1e-15, atol=0)
    # eigen systems may differ by ordering/sign; P(t) must agree
    S = 4
    for t in (0.0, 1e-4, 0.05, 0.7, 25.0):
        for r in got["rates"]:
            V, Vi = got["eigenvecs"].reshape(S, S), got["inv_eigenvecs"].reshape(S, S)
            P = np.eye(S) + (Vi * np.expm1(got["eigenvals"] * r * t)[None, :]) @ V
            Vw, Viw = want.eigenvecs.reshape(S, S), want.inv_eigenvecs.reshape(S, S)
            Pw = np.eye(S) + (Viw * np.expm1(want.eigenvals * r * t)[None, :]) @ Vw
            assert np.allclose(P, Pw, rtol=0, atol=1e-14)
            assert np.allclose(P.sum(axis=1), 1.0, atol=1e-13)


@pytest.mark.parametrize("model", ["LG+G4{0.8}", "WAG", "JTT+G4{1.3}", "DAYHOFF+FE+G2{0.5}"])
def test_protein_model_matches_oracle(sess, model):
    o = helpers.oracle()
    want = o.parse_model(model)
    got = sess.parse_model(model)
    S = 20
    assert got["states"] == S and got["rate_cats"] == want.rate_cats
    assert np.allclose(got["rates"], want.rates, rtol=1e-13, atol=0)
    assert np.allclose(got["freqs"], want.freqs, rtol=1e-15, atol=0)
    V, Vi = got["eigenvecs"].reshape(S, S), got["inv_eigenvecs"].reshape(S, S)
    Vw, Viw = want.eigenvecs.reshape(S, S), want.inv_eigenvecs.reshape(S, S)
    for t in (1e-3, 0.1, 2.0):
        P = np.eye(S) + (Vi * np.expm1(got["eigenvals"] * t)[None, :]) @ V
        Pw = np.eye(S) + (Viw * np.expm1(want.eigenvals * t)[None, :]) @ Vw
        assert np.allclose(P, Pw, rtol=0, atol=1e-13)


def test_model_errors(sess, built):
    for bad in ["GTR+IX", "GTR+IU{1.5}", "GTR+IU", "LG{1/2}+G", "GTR{1/2/3}", "GTR+FU{0.5/0.5}", "GTR+FQ", "GTR+R4{1/2}", "GTR+R4{1/2/3/4}{1/2}", "GTR+ASC_LEWIS", "FOO"]:
        with pytest.raises(built.capi.EpaError):
            sess.parse_model(bad)
    # +I in ML mode stays at the reference's unoptimised 0 (src/core/raxml/Model.cpp:192,355-380); +IU{p} is a user value
    for good in ["GTR+I+G4", "GTR+IO", "GTR+IC", "GTR+IU{0.2}+G4{0.5}", "LG+IU{0.1}+G4", "GTR+F+G4", "LG+FC+G4", "GTR+R4", "GTR+R2{0.3/2.0}", "LG+R4{0.1/0.5/1.2/3.0}{0.4/0.3/0.2/0.1}"]:
        sess.parse_model(good)


def test_schedule_reproduces_oracle_clvs(sess):
    """Runs the host layer's pruning schedule with the ORACLE's CLV kernel on the CPU and checks the
    reference's invariant (same tree log-likelihood across every edge, test/src/epa_pll_util.cpp:82-121)
    plus equality with the oracle's own directional CLVs."""
    o = helpers.oracle()
    case = helpers.synth64_case()
    n_slots, ops, edges, labels = sess.tree_schedule(_read("synth64", "tree.nwk"))
    T = len(labels)
    assert T == 64 and n_slots == 3 * (T - 2) and len(ops) == n_slots and len(edges) == 2 * T - 3
    by_label = {t.label: case.ref.sides[t.uid] for t in case.tree.tips}
    sides = {i: by_label[l] for i, l in enumerate(labels)}
    pending = list(ops)
    mc = case.model.c()
    while pending:
        rest = []
        for (p, l, r, ll, rl) in pending:
            if l in sides and r in sides:
                clv = np.zeros(case.n * 16)
                sc = np.zeros(case.n, dtype=np.uint32)
                o.lib().orc_update_partial(C.byref(mc), case.n, clv.ctypes.data_as(C.POINTER(C.c_double)),
                                           sc.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(sides[l].c),
                                           case.ref.pmat(ll).ctypes.data_as(C.POINTER(C.c_double)), C.byref(sides[r].c),
                                           case.ref.pmat(rl).ctypes.data_as(C.POINTER(C.c_double)))
                sides[p] = o.SideData(clv=clv, scaler=sc)
            else:
                rest.append((p, l, r, ll, rl))
        assert len(rest) < len(pending), "schedule has unsatisfiable dependencies"
        pending = rest
    want = case.ref.tree_logl(0)
    for e, (d, p, length) in enumerate(edges):
        assert d < T or p >= T
        got = o.lib().orc_edge_logl(C.byref(mc), case.n, C.byref(sides[d].c), C.byref(sides[p].c),
                                    case.ref.pmat(length).ctypes.data_as(C.POINTER(C.c_double)), None)
        assert abs(got - want) <= 1e-10 * abs(want), f"edge {e}"
        # same edge numbering and orientation as the oracle / reference
        od, op_, ol = case.ref.edges[e]
        assert ol == length
        if od.clv is not None:
            assert np.allclose(sides[d].clv, od.clv, rtol=1e-12, atol=0)
        assert np.allclose(sides[p].clv, op_.clv, rtol=1e-12, atol=0)


ROOTED = ["ref_rooted.tre", "ref_rooted_2.tre", "ref_rooted_3.tre", "ref_rooted_innerlabels.tre"]


@pytest.mark.parametrize("fname", ROOTED)
def test_rooted_numbered_newick_matches_reference(sess, fname):
    gold = __import__("json").load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_rooted.json")))[fname]
    text = _read("cfg1", fname)
    nwk, nt, ne = sess.parse_tree(text)
    assert (nt, ne) == (6, 9)                      # the working tree is unrooted: 2T - 3 edges
    assert nwk == gold["default"]["tree"]           # rooted tree, rooted edge numbers (10 of them)
    _, _, unrooted = sess.map_rooted(text, [0], [0.0])
    assert unrooted == gold["unrooted"]["tree"]     # --preserve-rooting off


def test_rtree_mapper_goldens(sess):
    """The reference's own known-answer vectors, test/src/rtree_mapper.cpp:58-102."""
    cases = {
        "ref_rooted.tre": ([(8, 1.0), (8, 1.5), (6, 0.5), (7, 0.001)], [(9, 1.0), (6, 0.63), (7, 0.5), (8, 0.001)]),
        "ref_rooted_2.tre": ([(0, 1.34), (0, 1.345), (8, 0.5), (2, 0.001)], [(0, 1.34), (9, 0.005), (8, 0.5), (2, 0.001)]),
        "ref_rooted_3.tre": ([(8, 0.5), (8, 0.005), (0, 0.5), (2, 0.001)], [(8, 1.41), (9, 0.005), (0, 0.5), (2, 0.001)]),
    }
    for fname, (u, r) in cases.items():
        e, d, _ = sess.map_rooted(_read("cfg1", fname), [x[0] for x in u], [x[1] for x in u])
        assert [int(x) for x in e] == [x[0] for x in r], fname
        assert np.allclose(d, [x[1] for x in r], atol=1e-10), fname


def test_bfast_reader_matches_fasta(built):
    """Query files in the reference's binary 4-bit format (src/io/Binary_Fasta.hpp, src/io/encoding.hpp):
    the fixtures were written by the reference's own converter (epa-ng --bfast)."""
    import os
    import numpy as np
    g = os.path.join(os.path.dirname(__file__), "golden")
    for fasta in (os.path.join(g, "bfast", "codes.fasta"), os.path.join(g, "cfg1", "query.fasta")):
        n1, r1 = built.session.read_alignment(fasta)
        n2, r2 = built.session.read_alignment(fasta + ".bfast")
        assert n1 == n2
        assert np.array_equal(r1, r2)
    # every 4-bit code occurs in the first fixture
    _, codes = built.session.read_alignment(os.path.join(g, "bfast", "codes.fasta.bfast"))
    assert set(bytes(codes.reshape(-1)).decode()) == set("-TGKCYSBAWRDMHVN")


def test_bfast_converter_is_byte_identical_to_the_reference(built, tmp_path):
    """-c/--bfast (Binary_Fasta::fasta_to_bfast, src/io/Binary_Fasta.hpp:214-246): the files written by the
    host layer equal, byte for byte, the ones the reference's converter wrote (committed fixtures);
    amino-acid input is refused like the reference does (ensure_dna)."""
    g = helpers.GOLDEN
    for fasta in (os.path.join(g, "bfast", "codes.fasta"), os.path.join(g, "cfg1", "query.fasta")):
        out = built.session.fasta_to_bfast(fasta, str(tmp_path))
        assert out == os.path.join(str(tmp_path), os.path.basename(fasta) + ".bfast")
        assert open(out, "rb").read() == open(fasta + ".bfast", "rb").read()
    with pytest.raises(built.capi.EpaError, match="AA DATA NOT SUPPORTED"):
        built.session.fasta_to_bfast(os.path.join(g, "synthaa", "query.fasta"), str(tmp_path))
    exe = os.path.join(helpers.ROOT, "epa-ng_b200", "epa-ng-b200")
    d = tmp_path / "cli"
    d.mkdir()
    import subprocess
    subprocess.run([exe, "-c", os.path.join(g, "bfast", "codes.fasta"), "-w", str(d)], check=True, stdout=subprocess.DEVNULL)
    assert open(d / "codes.fasta.bfast", "rb").read() == open(os.path.join(g, "bfast", "codes.fasta.bfast"), "rb").read()


def test_empirical_frequencies_match_oracle(built):
    """+F / +FC: the host layer counts the base frequencies on the reference MSA as the reference does
    (compute_and_set_empirical_frequencies); checked against the oracle, whose +F path is pinned on the
    reference's placements and printed frequencies (tests/test_oracle_freqs.py)."""
    for case, model in ((helpers.cfg1_case("GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FC+G4{1.0}"), "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FC+G4{1.0}"),
                        (helpers.load_case(*[os.path.join(helpers.GOLDEN, "synthaa", f) for f in ("tree.nwk", "ref.fasta", "query.fasta")],
                                           "LG+F+G4{0.8}"), "LG+F+G4{0.8}")):
        S = case.model.states
        f, ev = built.session.empirical_frequencies(model, case.tip_masks())
        assert np.allclose(f[:S], case.model.freqs, rtol=1e-13, atol=0)
        assert np.allclose(np.sort(ev[:S]), np.sort(case.model.eigenvals), rtol=1e-9, atol=1e-12)
    g = json.load(open(os.path.join(helpers.GOLDEN, "cfg1", "reference_empirical.json")))
    case = helpers.cfg1_case("GTR+FC+G4")
    f, _ = built.session.empirical_frequencies("GTR+FC+G4", case.tip_masks())
    assert np.allclose(f[:4], g["cfg1_printed_freqs"], atol=1e-6)


def test_model_files_give_the_reference_strings(built):
    """-m <file>: the reference's own golden strings (test/src/parse_model.cpp:7-70) on its own data files
    (test/data/modelfiles, copied as fixtures; the IQ-TREE report cut after its model section)."""
    d = os.path.join(helpers.GOLDEN, "modelfiles")
    want = {
        "rax8_dna": "GTR{0.787874/1.821672/1.294006/0.698421/3.034135/1.000000}+FU{0.256465/0.222535/0.308594/0.212406}+G4{0.478218}",
        "rax8_invar": "GTR{1.217620/2.720208/1.342850/1.115245/3.313319/1.000000}+FU{0.222438/0.209333/0.259930/0.308299}"
                      "+IU{0.051355}+G4{0.532224}",
        "raxng_dna": "GTR{5.56435/19.04/4.65971/2.04432/69.6551/1}+FC+G4m{0.193259}",
        "iqtree_dna_invar": "GTR{0.9467/3.2100/1.8644/0.8054/5.5442/1.0000}+FU{0.2415/0.2465/0.3237/0.1884}+IU{0.1257}+G4{0.8042}",
    }
    for f, s in want.items():
        assert built.session.model_from_file(os.path.join(d, f)) == s
        built.session.parse_model(s)                      # and the strings are models the host layer accepts
    prot = built.session.model_from_file(os.path.join(d, "rax8_prot"))
    assert prot.startswith("PROTGTR{1.003440/0.000100/2.196009/") and prot.endswith("}+G4{0.563473}")
    assert prot.count("/") == 189 + 19 and "+FU{0.065149/0.054231/" in prot
    m = built.session.parse_model(prot)
    assert m["states"] == 20 and abs(m["freqs"].sum() - 1.0) < 1e-12
    with pytest.raises(built.capi.EpaError, match="seems wrong"):
        built.session.model_from_file(os.path.join(helpers.GOLDEN, "cfg1", "query.fasta"))


def test_jplace_writer_reproduces_the_reference_file(built, tmp_path):
    """The jplace text (src/io/jplace_util.cpp:20-86: field order edge, logl, lwr, DISTAL, PENDANT; layout; metadata)
    for the reference's own numbers: a jplace written by the unmodified reference (cfg1, --no-heur --filter-max 3
    --filter-min-lwr 0) is parsed, its placements go through the host layer's writer, and the bytes must be equal."""
    import json
    src = os.path.join(helpers.GOLDEN, "cfg1", "reference_result.jplace")
    text = open(src).read()
    doc = json.loads(text)
    names = [pq["n"][0] for pq in doc["placements"]]
    fmax = max(len(pq["p"]) for pq in doc["placements"])
    recs = np.zeros((len(names), fmax), dtype=built.capi.PLACEMENT_DTYPE)
    counts = np.zeros(len(names), dtype=np.uint32)
    for i, pq in enumerate(doc["placements"]):
        counts[i] = len(pq["p"])
        for k, (edge, logl, lwr, distal, pendant) in enumerate(pq["p"]):
            recs[i, k] = (edge, logl, lwr, pendant, distal)
    out = str(tmp_path / "out.jplace")
    built.session.write_jplace(out, doc["tree"], doc["metadata"]["invocation"], names, recs, counts, precision=10)
    assert open(out).read() == text


def test_cli_takes_a_model_file_and_fails_loudly_without_a_gpu(tmp_path):
    """The CLI resolves -m <file> (src/main.cpp:433-436) before anything touches the device; without a CUDA device the
    run ends with an error message and a non-zero exit code - there is no CPU path to fall back to."""
    import subprocess
    exe = os.path.join(helpers.ROOT, "epa-ng_b200", "epa-ng-b200")
    d = os.path.join(helpers.GOLDEN, "cfg1")
    r = subprocess.run([exe, "-t", os.path.join(d, "ref.tre"), "-s", os.path.join(d, "aln.fasta"), "-q", os.path.join(d, "query.fasta"),
                        "-m", os.path.join(helpers.GOLDEN, "modelfiles", "rax8_dna"), "-w", str(tmp_path), "--redo"],
                       capture_output=True, text=True)
    out = r.stdout + r.stderr
    assert "==> model GTR{0.787874/1.821672/1.294006/0.698421/3.034135/1.000000}+FU{0.256465/0.222535/0.308594/0.212406}+G4{0.478218}" in out
    if "no CUDA device" in out:
        assert r.returncode != 0 and "no CPU path" in out
        assert not os.path.exists(os.path.join(str(tmp_path), "epa_result.jplace")) or os.path.getsize(os.path.join(str(tmp_path), "epa_result.jplace")) == 0
    else:
        assert r.returncode == 0


def test_jplace_strings_are_escaped(built, tmp_path):
    """Query names, the tree and the invocation go into JSON strings: quotes, backslashes and control characters are
    escaped, the file stays valid JSON and the strings come back unchanged."""
    import json
    names = ['plain', 'with "quotes"', 'back\\slash', 'tab\there', 'mix "\\" end']
    recs = np.zeros((len(names), 2), dtype=built.capi.PLACEMENT_DTYPE)
    counts = np.ones(len(names), dtype=np.uint32)
    for i in range(len(names)):
        recs[i, 0] = (i, -100.0 - i, 1.0, 0.1, 0.2)
    tree = "('t \"a\"':0.1{0},B:0.2{1},C\\\\x:0.3{2});"
    inv = 'epa-ng-b200 --model "GTR+G" -w C:\\out'
    out = str(tmp_path / "esc.jplace")
    built.session.write_jplace(out, tree, inv, names, recs, counts, precision=6)
    doc = json.loads(open(out).read())
    assert [pq["n"][0] for pq in doc["placements"]] == names
    assert doc["tree"] == tree and doc["metadata"]["invocation"] == inv
    assert doc["placements"][1]["p"][0][:2] == [1, -101.0]


def test_read_alignment_fails_when_the_label_buffer_is_too_small(built, tmp_path):
    """Long FASTA headers must not be truncated silently: a label buffer that cannot hold them is an error."""
    import ctypes as C
    path = str(tmp_path / "long.fasta")
    names = ["sequence_%03d_" % i + "x" * 150 for i in range(40)]
    with open(path, "w") as fh:
        for nm in names:
            fh.write(">%s\nACGTACGT\n" % nm)
    got, rows = built.session.read_alignment(path)
    assert got == names and rows.shape == (40, 8)
    L = built.session.lib()
    n, sites = C.c_uint32(), C.c_uint32()
    out = np.zeros((40, 8), dtype=np.uint8)
    small = C.create_string_buffer(64 * 40)
    rc = L.epa_host_read_alignment(path.encode(), C.byref(n), C.byref(sites), out.ctypes.data, out.size, small, len(small))
    assert rc != 0 and b"label" in L.epa_host_last_error().lower()
