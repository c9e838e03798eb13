"""Deterministic synthetic datasets for the BASELINE.json configurations.

Harness code (bench.py, tests): random unrooted binary tree, a reference MSA evolved down the tree
under the same model that is handed to the placement code, and window queries cut from mutated
reference tips. Pure numpy; independent of both the oracle and the CUDA library.
"""
from __future__ import annotations

import numpy as np

DNA = "ACGT"
AA = "ARNDCQEGHILKMFPSTWYV"


def random_tree(T: int, seed: int = 1, ladder: float = 0.0, brlen: float = 0.05):
    """Random topology by repeatedly joining two active nodes until 3 remain (trifurcating root);
    branch lengths Exp(mean brlen) + 0.001 printed with 6 decimals. With probability `ladder` a join
    takes the newest inner node and a random tip (deep, caterpillar-like trees: most CLV updates are
    tip-inner). Returns (newick, structure) where structure = (children, lengths, root_children) over
    node ids, tips = 0..T-1."""
    rng = np.random.default_rng(seed)
    active = list(range(T))
    children = {}
    length = {}
    nxt = T
    while len(active) > 3:
        i, j = sorted(rng.choice(len(active), size=2, replace=False))
        if ladder > 0.0 and active[-1] >= T and rng.random() < ladder:
            tips = [k for k in range(len(active) - 1) if active[k] < T]
            if tips:
                i, j = int(rng.choice(tips)), len(active) - 1
        a, b = active[i], active[j]
        active.pop(j)
        active.pop(i)
        children[nxt] = (a, b)
        active.append(nxt)
        nxt += 1
    for node in range(nxt):
        length[node] = round(float(rng.exponential(brlen)) + 0.001, 6)

    def name(t):
        return "t%04d" % t if T <= 10000 else "t%06d" % t

    def nwk(node):
        # iterative to survive caterpillar-ish shapes
        out, stack = [], [(node, 0)]
        while stack:
            x, st = stack.pop()
            if x < T:
                out.append("%s:%.6f" % (name(x), length[x]))
                continue
            a, b = children[x]
            if st == 0:
                out.append("(")
                stack.append((x, 2))
                stack.append((b, 0))
                stack.append((x, 1))
                stack.append((a, 0))
            elif st == 1:
                out.append(",")
            else:
                out.append("):%.6f" % length[x])
        return "".join(out)

    newick = "(" + ",".join(nwk(a) for a in active) + ");"
    return newick, (children, length, list(active), T)


def gtr_model_matrices(subst, freqs, rates, S):
    """Eigen system of the reversible rate matrix normalised to mean rate 1 (numpy)."""
    Q = np.zeros((S, S))
    k = 0
    for i in range(S):
        for j in range(i + 1, S):
            Q[i, j] = subst[k] * freqs[j]
            Q[j, i] = subst[k] * freqs[i]
            k += 1
    Q -= np.diag(Q.sum(axis=1))
    Q /= -(freqs * np.diag(Q)).sum()
    sq = np.sqrt(freqs)
    A = (sq[:, None] * Q) / sq[None, :]
    A = 0.5 * (A + A.T)
    lam, U = np.linalg.eigh(A)
    return lam, U, sq


def pmatrix(lam, U, sq, t):
    E = (U * np.exp(lam * t)[None, :]) @ U.T
    P = E * (sq[None, :] / sq[:, None])
    P = np.clip(P, 0, None)
    return P / P.sum(axis=1, keepdims=True)


def evolve_msa(structure, n_sites, subst, freqs, cat_rates, seed=1, eig=None):
    """uint8[T][n_sites] state indices evolved from a uniform root sequence. eig = (eigenvals, V, Vinv)
    in libpll layout replaces the numpy eigen-decomposition (protein models)."""
    children, length, roots, T = structure
    S = len(freqs)
    rng = np.random.default_rng(seed + 1000)
    if eig is None:
        lam, U, sq = gtr_model_matrices(np.asarray(subst, float), np.asarray(freqs, float), cat_rates, S)
        pm = lambda t: pmatrix(lam, U, sq, t)
    else:
        ev, V, Vi = eig

        def pm(t):
            P = np.clip(np.eye(S) + (Vi * np.expm1(ev * t)[None, :]) @ V, 0, None)
            return P / P.sum(axis=1, keepdims=True)
    site_cat = rng.integers(0, len(cat_rates), size=n_sites)
    root_seq = rng.integers(0, S, size=n_sites).astype(np.uint8)
    out = np.zeros((T, n_sites), dtype=np.uint8)

    def descend(parent_seq, node):
        stack = [(parent_seq, node)]
        while stack:
            pseq, x = stack.pop()
            seq = np.empty(n_sites, dtype=np.uint8)
            u = rng.random(n_sites)
            for c, r in enumerate(cat_rates):
                sel = site_cat == c
                if not sel.any():
                    continue
                cum = np.cumsum(pm(length[x] * r), axis=1)
                rows = cum[pseq[sel]]
                seq[sel] = (u[sel][:, None] > rows).sum(axis=1).clip(0, S - 1)
            if x < T:
                out[x] = seq
            else:
                a, b = children[x]
                stack.append((seq, b))
                stack.append((seq, a))

    for r in roots:
        descend(root_seq, r)
    return out


def make_queries(msa_states, n_queries, window, alphabet, seed=2, mut=0.05):
    """uint8[n_queries][n_sites] ASCII rows: a mutated copy of a random tip, kept only inside a
    contiguous window of `window` columns ('-' elsewhere). window >= n_sites keeps everything."""
    T, n = msa_states.shape
    S = len(alphabet)
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    tips = rng.integers(0, T, size=n_queries)
    w = min(window, n)
    starts = rng.integers(0, n - w + 1, size=n_queries)
    out = np.full((n_queries, n), ord('-'), dtype=np.uint8)
    CH = 65536
    for lo in range(0, n_queries, CH):
        hi = min(lo + CH, n_queries)
        m = hi - lo
        cols = starts[lo:hi, None] + np.arange(w)[None, :]
        st = msa_states[tips[lo:hi, None], cols]
        mutate = rng.random((m, w)) < mut
        st = np.where(mutate, rng.integers(0, S, size=(m, w)), st)
        np.put_along_axis(out[lo:hi], cols, letters[st], axis=1)
    return out


def dataset(T=1000, n_sites=1000, n_queries=1000, window=200, kind="dna", seed_tree=1, seed_q=2,
            alpha=None, ladder=0.0, brlen=0.05):
    """Returns dict(newick, names, ref (uint8 ASCII [T][n]), queries (uint8 ASCII [Q][n]), qnames,
    model string). DNA: GTR{1/1/1/1/1/1}+FU{.25/.25/.25/.25}+G4{0.5}; AA: LG+G4{0.8}."""
    from math import isfinite  # noqa: F401
    newick, st = random_tree(T, seed_tree, ladder, brlen)
    if kind == "dna":
        alphabet, S = DNA, 4
        subst, freqs = np.ones(6), np.full(4, 0.25)
        alpha = 0.5 if alpha is None else alpha
        model = "GTR{1/1/1/1/1/1}+FU{0.25/0.25/0.25/0.25}+G4{%g}" % alpha
        eig = None
    else:
        # amino acids: LG, eigen system from the host library's model parser (no device needed)
        from . import session
        alphabet, S = AA, 20
        alpha = 0.8 if alpha is None else alpha
        model = "LG+G4{%g}" % alpha
        pm = session.parse_model(model)
        subst, freqs = None, pm["freqs"]
        eig = (pm["eigenvals"], pm["eigenvecs"].reshape(S, S), pm["inv_eigenvecs"].reshape(S, S))
    cat_rates = discrete_gamma_mean(alpha, 4)
    states = evolve_msa(st, n_sites, subst, freqs, cat_rates, seed_tree, eig=eig)
    letters = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    ref = letters[states]
    names = [("t%04d" % t if T <= 10000 else "t%06d" % t) for t in range(T)]
    queries = make_queries(states, n_queries, window, alphabet, seed_q)
    qnames = ["q%07d" % i for i in range(n_queries)]
    return dict(newick=newick, names=names, ref=ref, queries=queries, qnames=qnames, model=model,
                states=S)


def discrete_gamma_mean(alpha, k):
    """Mean-of-category discrete gamma rates (generator only needs approximate rates)."""
    from scipy.stats import gamma as G
    from scipy.special import gammainc
    cuts = G.ppf(np.arange(1, k) / k, alpha, scale=1.0 / alpha)
    inc = np.concatenate([[0.0], gammainc(alpha + 1, cuts * alpha), [1.0]])
    return (inc[1:] - inc[:-1]) * k


def write_fasta(path, names, rows):
    rows = np.asarray(rows)
    if len(names) > 1000 and rows.ndim == 2 and len({len(n) for n in names[:1000]}) == 1 and len(names[0]) == len(names[-1]):
        # equal-length names: one 2-D byte array, written in one call
        L, n, w = len(names[0]), rows.shape[0], rows.shape[1]
        try:
            nm = np.frombuffer("".join(names).encode(), dtype=np.uint8).reshape(n, L)
        except ValueError:
            nm = None
        if nm is not None:
            out = np.empty((n, L + w + 3), dtype=np.uint8)
            out[:, 0] = ord(">")
            out[:, 1:1 + L] = nm
            out[:, 1 + L] = ord("\n")
            out[:, 2 + L:2 + L + w] = rows
            out[:, 2 + L + w] = ord("\n")
            out.tofile(path)
            return
    with open(path, "wb") as fh:
        for nm, row in zip(names, rows):
            fh.write(b">" + nm.encode() + b"\n")
            fh.write(bytes(row) + b"\n")


def write_dataset(ds, outdir):
    import os
    os.makedirs(outdir, exist_ok=True)
    open(os.path.join(outdir, "tree.nwk"), "w").write(ds["newick"] + "\n")
    write_fasta(os.path.join(outdir, "ref.fasta"), ds["names"], ds["ref"])
    write_fasta(os.path.join(outdir, "query.fasta"), ds["qnames"], ds["queries"])
    return (os.path.join(outdir, "tree.nwk"), os.path.join(outdir, "ref.fasta"),
            os.path.join(outdir, "query.fasta"))
