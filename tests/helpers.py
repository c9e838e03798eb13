"""Test helpers: builds the same placement problem twice - as a CPU oracle (oracle/pyoracle.py)
and as a libepa_b200 context - so that tests compare the two stage by stage."""
from __future__ import annotations

import json
import os
import sys
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
GTRG = "GTR{0.5/0.5/0.5/0.5/0.5/1.0}+FU{0.25/0.25/0.25/0.25}+G4{1.0}"
GTR_B = ("GTR{0.676278/2.012275/0.478487/0.753965/2.406436/1.0}"
        "+FU{0.245629/0.235012/0.253054/0.266305}+G4{1.078763}")


def pkg():
    return ge.load_package()


def oracle():
    o = ge.load_oracle()
    o.build()
    return o


@dataclass
class Case:
    model: object            # pyoracle.Model
    tree: object             # pyoracle.Tree
    ref: object              # pyoracle.Reference (oracle CLVs for every direction)
    placer: object           # pyoracle.Placer
    qnames: list
    qseqs: list              # masked query strings
    n: int

    @property
    def query_rows(self):
        return np.frombuffer("".join(self.qseqs).encode(), dtype=np.uint8).reshape(len(self.qseqs), self.n)

    # ---- node numbering handed to the library: tips 0..T-1, inner directions T.. ----
    def node_ids(self):
        ids, tips, inner = {}, 0, 0
        T = len(self.tree.tips)
        for t in self.tree.tips:
            ids[t.uid] = tips
            tips += 1
        for e in self.tree.branches:
            for x in (e, e.back):
                if x.next is not None and x.uid not in ids:
                    ids[x.uid] = T + inner
                    inner += 1
        return ids, T, inner

    def tip_masks(self):
        return np.stack([self.ref.sides[t.uid].tip for t in self.tree.tips]).astype(np.uint32)

    def ops(self, ids):
        out, seen = [], set()
        for e in self.tree.branches:
            for x in (e, e.back):
                if x.next is not None and x.uid not in seen:
                    seen.add(x.uid)
                    out.append((ids[x.uid], ids[x.next.back.uid], ids[x.next.next.back.uid],
                                x.next.length, x.next.next.length))
        return out

    def edges(self, ids, T):
        out = []
        for e in self.tree.branches:
            d, p = ids[e.uid], ids[e.back.uid]
            if p < T and d >= T:
                d, p = p, d                     # a tip is always distal (Tiny_Tree.cpp:64-74)
            out.append((d, p, e.length))
        return out


def load_case(tree_file, ref_msa, query_file, model_desc, premasking=True, opts=None) -> Case:
    o = oracle()
    model = o.parse_model(model_desc)
    rn, rs = o.read_fasta(ref_msa)
    qn, qs = o.read_fasta(query_file)
    if premasking:
        mask = o.gap_mask(rs) | o.gap_mask(qs)
        rs, qs = o.apply_mask(rs, mask), o.apply_mask(qs, mask)
    tree = o.build_tree(open(tree_file).read())
    ref = o.Reference(tree, model, rn, rs)
    placer = o.Placer(ref, opts or o.Options(premasking=premasking))
    return Case(model, tree, ref, placer, qn, qs, ref.n)


def case_from_arrays(newick, names, ref_rows, qnames, query_rows, model_desc, opts=None, per_rate=False,
                     bugcompat=False, column_mask=False) -> Case:
    """column_mask: drop the columns that are all-gap in the reference or in the queries, as the
    reference's pre-masking does before anything is computed (src/seq/MSA_Info.hpp:93-111)."""
    o = oracle()
    model = o.parse_model(model_desc)
    model.per_rate_scalers, model.bugcompat_focus = bool(per_rate), bool(bugcompat)
    rs = [bytes(r).decode() for r in ref_rows]
    qs = [bytes(r).decode() for r in query_rows]
    if column_mask:
        mask = o.gap_mask(rs) | o.gap_mask(qs)
        rs, qs = o.apply_mask(rs, mask), o.apply_mask(qs, mask)
    tree = o.build_tree(newick)
    ref = o.Reference(tree, model, list(names), rs)
    placer = o.Placer(ref, opts or o.Options())
    return Case(model, tree, ref, placer, list(qnames), qs, ref.n)


def make_context(case: Case, device=0, compute=True):
    """libepa_b200 context for the case; CLVs are computed ON THE DEVICE from the op list."""
    capi = pkg().capi
    ids, T, n_inner = case.node_ids()
    m = case.model
    ctx = capi.Context(states=m.states, rate_cats=m.rate_cats, sites=case.n, eigenvals=m.eigenvals,
                       eigenvecs=m.eigenvecs, inv_eigenvecs=m.inv_eigenvecs, freqs=m.freqs, rates=m.rates,
                       weights=m.weights, tip_masks=case.tip_masks(), n_clv_slots=n_inner,
                       edges=case.edges(ids, T), device=device, pinv=float(getattr(m, "pinv", 0.0)),
                       flags=(capi.EPA_FLAG_RATE_SCALERS if getattr(m, "per_rate_scalers", False) else 0)
                       | (capi.EPA_FLAG_BUGCOMPAT_FOCUS if getattr(m, "bugcompat_focus", False) else 0))
    ctx.ids = ids
    if compute:
        ctx.compute_clvs(case.ops(ids))
    return ctx


def golden(name):
    return json.load(open(os.path.join(GOLDEN, name, "reference_placements.json")))


def cfg1_case(model=GTRG, **kw) -> Case:
    d = os.path.join(GOLDEN, "cfg1")
    return load_case(os.path.join(d, "ref.tre"), os.path.join(d, "aln.fasta"), os.path.join(d, "query.fasta"), model, **kw)


def synth64_case(query_file="query.fasta", **kw) -> Case:
    d = os.path.join(GOLDEN, "synth64")
    model = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{0.5}"
    return load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, query_file), model, **kw)


def synthaa_case(query_file="query.fasta", **kw) -> Case:
    d = os.path.join(GOLDEN, "synthaa")
    return load_case(os.path.join(d, "tree.nwk"), os.path.join(d, "ref.fasta"), os.path.join(d, query_file),
                     "LG+G4{0.8}", **kw)


def assert_placements_close(got, want, what="", logl_rel=1e-6, lwr_abs=1e-6, len_abs=1e-4):
    """got/want: lists of (edge, logl, lwr, distal, pendant) sorted by LWR descending."""
    assert [int(g[0]) for g in got] == [int(w[0]) for w in want], f"{what}: edge lists differ {got} vs {want}"
    for g, w in zip(got, want):
        assert abs(g[1] - w[1]) <= logl_rel * abs(w[1]), f"{what}: logl {g} vs {w}"
        assert abs(g[2] - w[2]) <= lwr_abs, f"{what}: lwr {g} vs {w}"
        assert abs(g[3] - w[3]) <= len_abs, f"{what}: distal {g} vs {w}"
        assert abs(g[4] - w[4]) <= len_abs, f"{what}: pendant {g} vs {w}"


def records_to_lists(out, counts):
    res = []
    for row, c in zip(out, counts):
        res.append([(int(r["branch_id"]), float(r["likelihood"]), float(r["lwr"]), float(r["distal_length"]),
                     float(r["pendant_length"])) for r in row[:c]])
    return res
