"""oracle/pyoracle.py - TEST INFRASTRUCTURE ONLY.

Python driver around oracle/epa_oracle.c: builds the C restatement, exposes it through ctypes and
re-states the host-side orchestration of EPA-ng (tree numbering, masking, chunk pipeline) so that a
complete placement run can be produced on the CPU and compared with (a) the unmodified reference
binary oracle/_ref/epa-ng and (b) the CUDA product. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this module; the product never does.

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
import subprocess
import sys
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# --------------------------------------------------------------------------------------------
#  build + ctypes
# --------------------------------------------------------------------------------------------
def lib_path() -> str:
    return os.path.join(HERE, "_build", "libepa_oracle.so")


def build(force: bool = False) -> str:
    """Compile epa_oracle.c -> oracle/_build/libepa_oracle.so (gcc, -O2, strict IEEE)."""
    src = os.path.join(HERE, "epa_oracle.c")
    hdr = os.path.join(HERE, "epa_oracle.h")
    out = lib_path()
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-o", out, src, "-lm"]
    subprocess.check_call(cmd)
    return out


class OrcModel(C.Structure):
    _fields_ = [("states", C.c_int), ("rate_cats", C.c_int), ("per_rate_scalers", C.c_int),
                ("bugcompat_focus", C.c_int),
                ("eigenvals", C.POINTER(C.c_double)), ("eigenvecs", C.POINTER(C.c_double)),
                ("inv_eigenvecs", C.POINTER(C.c_double)), ("freqs", C.POINTER(C.c_double)),
                ("rates", C.POINTER(C.c_double)), ("weights", C.POINTER(C.c_double)),
                ("pinv", C.c_double), ("invariant", C.POINTER(C.c_int))]


class OrcSide(C.Structure):
    _fields_ = [("clv", C.POINTER(C.c_double)), ("scaler", C.POINTER(C.c_uint32)),
                ("tip", C.POINTER(C.c_uint32))]


class OrcBlo(C.Structure):
    _fields_ = [("logl", C.c_double), ("pendant", C.c_double), ("distal", C.c_double),
                ("rounds", C.c_int), ("restored", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, up, u8p, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_int)
        mp, sp = C.POINTER(OrcModel), C.POINTER(OrcSide)
        L.orc_gamma_rates.argtypes = [C.c_double, C.c_int, C.c_int, dp]
        L.orc_gamma_rates.restype = C.c_int
        L.orc_eigen.argtypes = [C.c_int, dp, dp, dp, dp, dp]
        L.orc_eigen.restype = C.c_int
        L.orc_pmatrix.argtypes = [mp, C.c_double, dp]
        L.orc_invariant_sites.argtypes = [C.c_int, C.c_int, C.c_int, up, ip]
        L.orc_update_partial.argtypes = [mp, C.c_int, dp, up, sp, dp, sp, dp]
        L.orc_edge_logl.argtypes = [mp, C.c_int, sp, sp, dp, dp]
        L.orc_edge_logl.restype = C.c_double
        L.orc_sumtable.argtypes = [mp, C.c_int, sp, sp, dp]
        L.orc_derivatives.argtypes = [mp, C.c_int, dp, C.c_double, dp, dp]
        L.orc_tiny_inner.argtypes = [mp, C.c_int, sp, sp, C.c_double, dp, up]
        L.orc_lookup_build.argtypes = [mp, C.c_int, sp, sp, C.c_double, up, C.c_int, dp]
        L.orc_preplace_score.argtypes = [dp, C.c_int, u8p, C.c_int, C.c_int]
        L.orc_preplace_score.restype = C.c_double
        L.orc_place_thorough.argtypes = [mp, C.c_int, sp, sp, C.c_double, up, C.c_int, C.c_int,
                                         C.POINTER(OrcBlo)]
        L.orc_place_thorough_raxml.argtypes = L.orc_place_thorough.argtypes
        L.orc_lwr.argtypes = [dp, C.c_int, dp]
        L.orc_select_accumulated.argtypes = [dp, C.c_int, C.c_double, ip]
        L.orc_select_accumulated.restype = C.c_int
        L.orc_filter_support.argtypes = [dp, C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_filter_support.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


# --------------------------------------------------------------------------------------------
#  character maps (libs/pll-modules/libs/libpll/src/maps.c:46-111, src/util/maps.hpp:9-28,
#  src/core/Lookup_Store.hpp:33-68)
# --------------------------------------------------------------------------------------------
NT_MAP = "-TGKCYSBAWRDMHVN"
AA_MAP = "ACDEFGHIKLMNPQRSTVWY-XBZ"

_PLL_NT = {'A': 1, 'B': 14, 'C': 2, 'D': 13, 'G': 4, 'H': 11, 'K': 12, 'M': 3, 'N': 15, 'O': 15,
           'R': 5, 'S': 6, 'T': 8, 'U': 8, 'V': 7, 'W': 9, 'X': 15, 'Y': 10, '-': 15, '.': 15, '?': 15}
_AA_ORDER = "ARNDCQEGHILKMFPSTWYV"
_PLL_AA = {c: 1 << i for i, c in enumerate(_AA_ORDER)}
_PLL_AA.update({'B': _PLL_AA['N'] | _PLL_AA['D'], 'Z': _PLL_AA['Q'] | _PLL_AA['E'],
                'J': _PLL_AA['I'] | _PLL_AA['L'], 'X': 0xfffff, '*': 0xfffff, '-': 0xfffff,
                '.': 0xfffff, '?': 0xfffff})


def state_mask_table(states: int) -> np.ndarray:
    """uint32[256]: ASCII -> libpll state mask (0 = invalid), case-insensitive."""
    t = np.zeros(256, dtype=np.uint32)
    src = _PLL_NT if states == 4 else _PLL_AA
    for c, m in src.items():
        t[ord(c)] = m
        t[ord(c.lower())] = m
    return t


def lookup_char_map(states: int) -> str:
    return NT_MAP if states == 4 else AA_MAP


def lookup_column_table(states: int) -> np.ndarray:
    """uint8[256]: ASCII -> column of the preplacement lookup (255 = invalid)."""
    cmap = lookup_char_map(states)
    t = np.full(256, 255, dtype=np.uint8)
    for i, c in enumerate(cmap):
        t[ord(c)] = i
        t[ord(c.lower())] = i
    if states == 4:
        for c in "Uu":
            t[ord(c)] = t[ord('T')]
        for c in "XxOo.":
            t[ord(c)] = t[ord('-')]
    else:
        for c in "Xx":
            t[ord(c)] = t[ord('N')]          # reference quirk 2
    t[ord('?')] = t[ord('-')]
    return t


def lookup_masks(states: int) -> np.ndarray:
    tab = state_mask_table(states)
    return np.array([tab[ord(c)] for c in lookup_char_map(states)], dtype=np.uint32)


# --------------------------------------------------------------------------------------------
#  model
# --------------------------------------------------------------------------------------------
@dataclass
class Model:
    states: int
    subst: np.ndarray
    freqs: np.ndarray
    alpha: float
    rate_cats: int
    rates: np.ndarray
    weights: np.ndarray
    eigenvals: np.ndarray = None
    eigenvecs: np.ndarray = None
    inv_eigenvecs: np.ndarray = None
    per_rate_scalers: bool = False
    bugcompat_focus: bool = False
    pinv: float = 0.0                # +IU{p} (src/core/raxml/Model.cpp:355-380)
    empirical_freqs: bool = False    # +F / +FC: frequencies counted on the reference MSA (filled by Reference)
    invariant: np.ndarray = None     # int32[n], filled by Reference from the tip masks when pinv > 0
    _c: OrcModel = field(default=None, repr=False)

    def finalize(self):
        S = self.states
        self.eigenvals = np.zeros(S)
        self.eigenvecs = np.zeros(S * S)
        self.inv_eigenvecs = np.zeros(S * S)
        ok = lib().orc_eigen(S, _dp(self.subst), _dp(self.freqs), _dp(self.eigenvals),
                             _dp(self.eigenvecs), _dp(self.inv_eigenvecs))
        assert ok
        return self

    def c(self) -> OrcModel:
        m = OrcModel(self.states, self.rate_cats, int(self.per_rate_scalers), int(self.bugcompat_focus),
                     _dp(self.eigenvals), _dp(self.eigenvecs), _dp(self.inv_eigenvecs),
                     _dp(self.freqs), _dp(self.rates), _dp(self.weights), float(self.pinv),
                     self.invariant.ctypes.data_as(C.POINTER(C.c_int)) if self.invariant is not None else None)
        self._c = m
        return m

    def pmatrix(self, t: float) -> np.ndarray:
        out = np.zeros(self.rate_cats * self.states * self.states)
        lib().orc_pmatrix(C.byref(self.c()), t, _dp(out))
        return out


_PROT = None


def protein_tables():
    global _PROT
    if _PROT is None:
        _PROT = json.load(open(os.path.join(HERE, "protein_models.json")))
    return _PROT


def _read_braces(s: str, pos: int):
    """Parses an optional {a/b/c} at s[pos:]; returns (values|None, new_pos)."""
    if pos < len(s) and s[pos] == '{':
        end = s.index('}', pos)
        return [float(x) for x in s[pos + 1:end].split('/')], end + 1
    return None, pos


def gamma_rates(alpha: float, ncat: int, median: bool = False) -> np.ndarray:
    out = np.zeros(ncat)
    assert lib().orc_gamma_rates(alpha, ncat, int(median), _dp(out))
    return out


# rate symmetries AC AG AT CG CT GT of the named DNA models and their aliases (PM/util/models_dna.c:40-125)
_DNA_SYM = {"JC": [0] * 6, "F81": [0] * 6, "K80": [0, 1, 0, 0, 1, 0], "HKY": [0, 1, 0, 0, 1, 0],
            "TN93EF": [0, 1, 0, 0, 2, 0], "TN93": [0, 1, 0, 0, 2, 0], "K81": [0, 1, 2, 2, 1, 0], "K81UF": [0, 1, 2, 2, 1, 0],
            "TPM2": [0, 1, 0, 2, 1, 2], "TPM2UF": [0, 1, 0, 2, 1, 2], "TPM3": [0, 1, 2, 0, 1, 2], "TPM3UF": [0, 1, 2, 0, 1, 2],
            "TIM1": [0, 1, 2, 2, 3, 0], "TIM1UF": [0, 1, 2, 2, 3, 0], "TIM2": [0, 1, 0, 2, 3, 2], "TIM2UF": [0, 1, 0, 2, 3, 2],
            "TIM3": [0, 1, 2, 0, 3, 2], "TIM3UF": [0, 1, 2, 0, 3, 2], "TVMEF": [0, 1, 2, 3, 1, 4], "TVM": [0, 1, 2, 3, 1, 4],
            "SYM": list(range(6)), "GTR": list(range(6))}
_DNA_ALIASES = {"TRNEF": "TN93EF", "TRN": "TN93", "TPM1": "K81", "TPM1UF": "K81UF", "TPM2EF": "TPM2", "TPM3EF": "TPM3",
                "TIM1EF": "TIM1", "TIM2EF": "TIM2", "TIM3EF": "TIM3"}


def parse_model(desc: str) -> Model:
    """Subset of the raxml-ng model grammar of src/core/raxml/Model.cpp:123-560:
    DNA: the 22 named models of PM/util/models_dna.c (JC ... GTR) and their aliases, with optional {rates}; +FU{..}/+FE/+FO/+F; +G[n][a|m][{alpha}];
    +F / +FC (empirical, counted on the reference MSA by Reference); +IU{p} (+I / +IO / +IC stay at 0 as
    in the reference, which neither optimises nor counts the value).
    +R[n]{rates}{weights} (free rates). (ASC is outside the oracle's scope.)"""
    pos = len(desc)
    for ch in "+{[":
        p = desc.find(ch)
        if p != -1:
            pos = min(pos, p)
    name, opts = desc[:pos].upper(), desc[pos:]
    prot = protein_tables()
    name = _DNA_ALIASES.get(name, name)
    if name not in _DNA_SYM and name not in ("DNA", "PROTGTR") and name not in prot:
        raise ValueError(f"oracle: unsupported model name {name}")
    if name == "DNA":
        name, opts = "GTR", "+G+FO"
    if name == "PROTGTR":
        # protein GTR (PM/util/models_aa.c:69): 190 user exchangeabilities, ML-mode defaults otherwise
        S = 20
        sym, nuniq = None, 190
        subst = np.array([0.5] * 189 + [1.0])
        freqs = np.full(S, 1.0 / S)
    elif name in prot:
        # empirical protein matrix: exchangeabilities and frequencies of the model
        # (libs/pll-modules/libs/libpll/src/maps.c:288-, :1472-; data in oracle/protein_models.json)
        S = 20
        sym, nuniq = None, 0
        subst = np.array(prot[name]["rates"], dtype=float)
        freqs = np.array(prot[name]["freqs"], dtype=float)
    else:
        S = 4
        sym = _DNA_SYM[name]
        nuniq = max(sym) + 1
        # JC / F81 carry the model's equal rates; every other DNA model without {rates} starts from the ML-mode
        # default 0.5, 0.5, 0.5, 0.5, 0.5, 1.0 over the SIX rates, whatever its symmetry (Model.cpp:484-490)
        subst = np.ones(6) if name in ("JC", "F81") else np.array([0.5] * 5 + [1.0])
        freqs = np.full(S, 1.0 / S)
    alpha, ncat, median, gamma, pinv, empirical = 1.0, 1, False, False, 0.0, False
    free_rates = free_weights = None
    vals, i = _read_braces(opts, 0)
    if vals is not None:
        if sym is None and name != "PROTGTR":
            raise ValueError("user-defined rates need PROTGTR for protein data")
        if len(vals) != nuniq:
            raise ValueError("wrong number of substitution rates")
        if sym is None:
            subst = np.array(vals, dtype=float) / vals[-1]
        else:
            last = vals[sym[-1]]
            vals = [v / last for v in vals]
            subst = np.array([vals[c] for c in sym], dtype=float)
    while i < len(opts):
        ch = opts[i].upper()
        i += 1
        if ch == '+':
            continue
        if ch == 'F':
            mode = opts[i].upper() if i < len(opts) and opts[i] != '+' else 'C'
            if i < len(opts) and opts[i] != '+':
                i += 1
            if mode == 'U':
                vals, i = _read_braces(opts, i)
                f = np.array(vals, dtype=float)
                freqs = f / f.sum()
            elif mode in ('E', 'O'):
                freqs = np.full(S, 1.0 / S)
            elif mode == 'C':
                empirical = True             # counted by Reference once the tips are known
            else:
                raise ValueError("Invalid frequencies specification")
        elif ch == 'I':
            mode = opts[i].upper() if i < len(opts) and opts[i] != '+' else 'O'
            if i < len(opts) and opts[i] != '+':
                i += 1
            if mode == 'U':
                vals, i = _read_braces(opts, i)
                if vals is None:
                    raise ValueError("Invalid p-inv specification")
                pinv = vals[0]
                if not (0.0 <= pinv < 1.0):
                    raise ValueError("Invalid proportion of invariant sites")
            elif mode == 'C':
                pinv = 0.0        # the reference never computes the empirical value: it stays 0 ("P-inv (empirical): 0")
            elif mode != 'O':
                raise ValueError("Invalid p-inv specification")
        elif ch == 'G':
            gamma = True
            num = ""
            while i < len(opts) and opts[i].isdigit():
                num += opts[i]
                i += 1
            ncat = int(num) if num else 4
            if i < len(opts) and opts[i] in "aA":
                median, i = True, i + 1
            elif i < len(opts) and opts[i] in "mM":
                i += 1
            vals, i = _read_braces(opts, i)
            if vals is not None:
                alpha = vals[0]
        elif ch == 'R':
            # free rates (Model.cpp:405-455): +R[n]{rates}{weights}; weights normalised to sum 1, rates to
            # mean 1; without values the categories start as GAMMA(alpha = 1) with equal weights
            gamma = True
            num = ""
            while i < len(opts) and opts[i].isdigit():
                num += opts[i]
                i += 1
            ncat = int(num) if num else (4 if ncat == 1 else ncat)
            vals, i = _read_braces(opts, i)
            if vals is not None:
                if len(vals) != ncat:
                    raise ValueError("Invalid number of free rates specified")
                free_rates = np.array(vals, dtype=float)
                vals, i = _read_braces(opts, i)
                if vals is not None:
                    if len(vals) != ncat:
                        raise ValueError("Invalid number of rate weights specified")
                    free_weights = np.array(vals, dtype=float) / float(np.sum(vals))
                else:
                    free_weights = np.full(ncat, 1.0 / ncat)
                free_rates = free_rates / float((free_rates * free_weights).sum())
        else:
            raise ValueError(f"oracle: unsupported model option +{ch}")
    rates = gamma_rates(alpha, ncat, median) if gamma and ncat > 1 else np.ones(ncat)
    weights = np.full(ncat, 1.0 / ncat)
    if free_rates is not None:
        rates, weights = free_rates, free_weights
    # pll_set_frequencies (LP/models.c:445-470): frequencies that do not sum to 1 within 1e-8 are normalised
    # (the published protein tables carry six digits: LG sums to 1.000001, WAG to 0.9999999)
    if abs(freqs.sum() - 1.0) > 1e-8:
        freqs = freqs / freqs.sum()
    return Model(S, subst, freqs, alpha, ncat, rates, weights, pinv=pinv, empirical_freqs=empirical).finalize()


# --------------------------------------------------------------------------------------------
#  FASTA + masking (src/seq/MSA_Info.hpp:22-111, src/seq/MSA_Stream.cpp:8-36)
# --------------------------------------------------------------------------------------------
def read_fasta(path: str):
    names, seqs, cur = [], [], []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if not line:
                continue
            if line[0] == '>':
                if names:
                    seqs.append("".join(cur))
                names.append(line[1:])
                cur = []
            else:
                cur.append(line.strip().upper())
    if names:
        seqs.append("".join(cur))
    return names, seqs


_GAP_CHARS = set("NOX.-?noxn")


def gap_mask(seqs) -> np.ndarray:
    """True where EVERY sequence has one of genesis' 'undetermined nucleic acid' chars
    (the reference uses that set for protein data too)."""
    arr = np.frombuffer("".join(seqs).encode(), dtype=np.uint8).reshape(len(seqs), -1)
    is_gap = np.zeros(256, dtype=bool)
    for c in _GAP_CHARS:
        is_gap[ord(c)] = True
    return is_gap[arr].all(axis=0)


def apply_mask(seqs, mask: np.ndarray):
    keep = ~mask
    out = []
    for s in seqs:
        a = np.frombuffer(s.encode(), dtype=np.uint8)
        out.append(a[keep].tobytes().decode())
    return out


def valid_range(seq: str):
    """src/util/Range.hpp:34-49: only '-' counts as gap here."""
    lo, hi = 0, len(seq)
    while lo < hi and seq[lo] == '-':
        lo += 1
    while hi > lo and seq[hi - 1] == '-':
        hi -= 1
    return lo, hi - lo


# --------------------------------------------------------------------------------------------
#  tree (libpll utree semantics: parse_utree.y:187-319, src/core/pll/pll_util.cpp:182-352)
# --------------------------------------------------------------------------------------------
class UNode:
    __slots__ = ("next", "back", "label", "length", "uid")

    def __init__(self):
        self.next = None
        self.back = None
        self.label = None
        self.length = 0.0
        self.uid = -1


def _tokenize(s: str):
    i, n = 0, len(s)
    while i < n:
        c = s[i]
        if c in " \t\r\n":
            i += 1
        elif c in "(),:;":
            yield c
            i += 1
        elif c in "'\"":
            j = s.index(c, i + 1)
            yield ("L", s[i + 1:j])
            i = j + 1
        else:
            j = i
            while j < n and s[j] not in " \t\r\n()[],:;":
                j += 1
            yield ("L", s[i:j])
            i = j


def parse_newick(text: str) -> UNode:
    """Returns the virtual root (an inner unode of the top-level trifurcation), built exactly like
    libpll's grammar actions do (child order = file order)."""
    toks = list(_tokenize(text))
    pos = 0

    def label_length():
        nonlocal pos
        label, length = None, None
        if pos < len(toks) and isinstance(toks[pos], tuple):
            label = toks[pos][1]
            pos += 1
        if pos < len(toks) and toks[pos] == ':':
            length = float(toks[pos + 1][1])
            pos += 2
        return label, length

    def subtree():
        nonlocal pos
        if toks[pos] == '(':
            items = desc_list()
            node = UNode()
            node.label, ln = label_length()
            node.length = ln if ln is not None else 0.0
            ring = [node] + items
            for a, b in zip(ring, ring[1:] + ring[:1]):
                a.next = b
            for it in items:
                if it.label is None:
                    it.label = node.label
            return node
        node = UNode()
        node.label, ln = label_length()
        node.length = ln if ln is not None else 0.0
        return node

    def desc_list():
        nonlocal pos
        assert toks[pos] == '('
        items = []
        while True:
            pos += 1
            sub = subtree()
            it = UNode()
            it.back, sub.back = sub, it
            it.length = sub.length
            items.append(it)
            if toks[pos] != ',':
                break
        assert toks[pos] == ')', "newick: expected ')'"
        pos += 1
        return items

    items = desc_list()
    label, _ = label_length()
    assert toks[pos] == ';'
    for a, b in zip(items, items[1:] + items[:1]):
        a.next = b
    for it in items:
        it.label = label
    return items[0]


def _ring_size(n: UNode) -> int:
    k, x = 1, n.next
    while x is not n:
        k += 1
        x = x.next
    return k


def unroot(root: UNode) -> UNode:
    """Rooted (bifurcating top level) input is outside the oracle's scope for now."""
    if _ring_size(root) != 3:
        raise ValueError("oracle: only unrooted (trifurcating) reference trees are supported")
    return root


DEFAULT_BRANCH_LENGTH = -math.log(0.9)


@dataclass
class Tree:
    root: UNode
    branches: list          # edge nodes in utree_query_branches order
    tips: list              # tip unodes
    num_sites: int = 0

    @property
    def num_branches(self):
        return len(self.branches)


def build_tree(newick_text: str) -> Tree:
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    root = unroot(parse_newick(newick_text))
    # set_missing_branch_lengths (pll_util.cpp:17-44)
    branches, tips = [], []

    def rec(node):
        if node.next is not None:
            assert _ring_size(node) == 3, "Input tree contains multifurcations"
            rec(node.next.back)
            rec(node.next.next.back)
        else:
            tips.append(node)
        branches.append(node)

    rec(root.back)
    rec(root.next.back)
    rec(root.next.next.back)
    for e in branches:
        if not e.length:
            e.length = DEFAULT_BRANCH_LENGTH
        e.back.length = e.length
    uid = 0
    for e in branches:
        for x in (e, e.back):
            if x.uid < 0:
                x.uid = uid
                uid += 1
    return Tree(root, branches, tips)


def numbered_newick(tree: Tree, precision: int = 10) -> str:
    """pll_util.cpp:207-352, unrooted case."""
    idx = [0]
    fmt = "%." + str(precision) + "f"

    def rec(node):
        if node.next is not None:
            s = "(" + rec(node.next.back) + "," + rec(node.next.next.back) + ")"
            if node.label:
                s += node.label
        else:
            s = node.label
        s += ":" + (fmt % node.length) + "{" + str(idx[0]) + "}"
        idx[0] += 1
        return s

    r = tree.root
    out = "(" + rec(r.back) + "," + rec(r.next.back) + "," + rec(r.next.next.back) + ")"
    if r.label:
        out += r.label
    return out + ";"


# --------------------------------------------------------------------------------------------
#  reference state: directional CLVs for every edge
# --------------------------------------------------------------------------------------------
class SideData:
    """Owns the numpy buffers behind an orc_side_t."""

    def __init__(self, clv=None, scaler=None, tip=None):
        self.clv, self.scaler, self.tip = clv, scaler, tip
        self.c = OrcSide(_dp(clv), _up(scaler), _up(tip))

    @property
    def is_tip(self):
        return self.tip is not None


class Reference:
    """Reference tree + MSA + model with all directional CLVs precomputed
    (src/tree/Tree.cpp:16-56, src/core/pll/epa_pll_util.cpp:10-107)."""

    def __init__(self, tree: Tree, model: Model, names, seqs):
        self.tree, self.model = tree, model
        self.n = len(seqs[0])
        S, R = model.states, model.rate_cats
        tab = state_mask_table(S)
        by_name = dict(zip(names, seqs))
        self.sides = {}
        for t in tree.tips:
            if t.label not in by_name:
                raise ValueError(f"taxon {t.label} missing from reference MSA")
            a = np.frombuffer(by_name[t.label].encode(), dtype=np.uint8)
            m = tab[a].astype(np.uint32)
            if (m == 0).any():
                raise ValueError("invalid character in reference MSA")
            self.sides[t.uid] = SideData(tip=np.ascontiguousarray(m))
        self._pm_cache = {}
        if model.empirical_freqs:
            # compute_and_set_empirical_frequencies (src/core/pll/optimize.cpp:457-472) ->
            # pllmod_msa_empirical_frequencies (PM/msa/pll_msa.c:45-143): every tip character spreads one
            # count evenly over the states of its mask; divided by sites * tips
            f = np.zeros(S)
            for t in tree.tips:
                m = self.sides[t.uid].tip
                pop = np.array([bin(int(x)).count("1") for x in m], dtype=float)
                for k in range(S):
                    f[k] += (((m >> k) & 1) / pop).sum()
            model.freqs = f / (self.n * len(tree.tips))
            model.finalize()
        if model.pinv > 0:
            # pll_update_invariant_sites_proportion -> pll_update_invariant_sites, called by
            # raxml::assign after the tips are linked (src/core/pll/epa_pll_util.cpp:59)
            masks = np.ascontiguousarray(np.stack([self.sides[t.uid].tip for t in tree.tips]))
            model.invariant = np.zeros(self.n, dtype=np.int32)
            lib().orc_invariant_sites(S, len(tree.tips), self.n, _up(masks),
                                      model.invariant.ctypes.data_as(C.POINTER(C.c_int)))
        mc = model.c()
        ssz = self.n * (R if model.per_rate_scalers else 1)

        def side(node):
            # iterative post-order to compute the CLV "behind" node
            stack = [(node, False)]
            while stack:
                x, ready = stack.pop()
                if x.uid in self.sides:
                    continue
                c1, c2 = x.next.back, x.next.next.back
                if not ready:
                    stack.append((x, True))
                    stack.append((c1, False))
                    stack.append((c2, False))
                    continue
                clv = np.zeros(self.n * R * S)
                sc = np.zeros(ssz, dtype=np.uint32)
                l, r = self.sides[c1.uid], self.sides[c2.uid]
                lib().orc_update_partial(C.byref(mc), self.n, _dp(clv), _up(sc),
                                         C.byref(l.c), _dp(self.pmat(x.next.length)),
                                         C.byref(r.c), _dp(self.pmat(x.next.next.length)))
                self.sides[x.uid] = SideData(clv=clv, scaler=sc)
            return self.sides[node.uid]

        self.edges = []     # (distal side, proximal side, length) with the tip always distal
        for e in tree.branches:
            d, p = side(e), side(e.back)
            if p.is_tip and not d.is_tip:
                d, p = p, d             # Tiny_Tree.cpp:64-74
            self.edges.append((d, p, e.length))

    def pmat(self, t: float):
        if t not in self._pm_cache:
            self._pm_cache[t] = self.model.pmatrix(t)
        return self._pm_cache[t]

    def tree_logl(self, edge: int = 0) -> float:
        """Reference-tree log-likelihood evaluated at an edge (Tree.cpp:119-131); equal on every
        edge (test/src/epa_pll_util.cpp:82-121)."""
        e = self.tree.branches[edge]
        a, b = self.sides[e.uid], self.sides[e.back.uid]
        pm = self.pmat(e.length)
        return lib().orc_edge_logl(C.byref(self.model.c()), self.n, C.byref(a.c), C.byref(b.c), _dp(pm), None)


# --------------------------------------------------------------------------------------------
#  placement pipeline (src/core/place.cpp:173-251)
# --------------------------------------------------------------------------------------------
@dataclass
class Options:
    prescoring: bool = True
    prescoring_threshold: float = 0.99999
    support_threshold: float = 0.01
    filter_min: int = 1
    filter_max: int = 7
    premasking: bool = True
    heuristic: int = 0          # 0 dynamic (-g), 1 fixed fraction (-G), 2 baseball
    sliding_blo: bool = True    # False = --raxml-blo (src/core/pll/optimize.cpp:274-278)


@dataclass
class Placement:
    edge: int
    logl: float
    lwr: float
    pendant: float
    distal: float


class Placer:
    def __init__(self, ref: Reference, opts: Options = None):
        self.ref, self.opts = ref, opts or Options()
        S = ref.model.states
        self.masks = lookup_masks(S)
        self.K = len(self.masks)
        self.col_tab = lookup_column_table(S)
        self.mask_tab = state_mask_table(S)
        self.lookup = None
        # Reference quirk: the tiny partition aliases the reference partition's category RATES but not its rate
        # WEIGHTS (src/tree/tiny_util.cpp:110-111), so every tiny-tree likelihood uses pll_partition_create's
        # default weights 1/R - visible only with user-defined free-rate weights (+R{..}{..}).
        m = ref.model
        self.pmodel = m
        if not np.allclose(m.weights, 1.0 / m.rate_cats, rtol=0, atol=0):
            import dataclasses
            self.pmodel = dataclasses.replace(m, weights=np.full(m.rate_cats, 1.0 / m.rate_cats), _c=None)

    def build_lookup(self):
        ref, mc = self.ref, self.pmodel.c()
        B, n, K = len(ref.edges), ref.n, self.K
        self.lookup = np.zeros((B, n, K))
        for b, (d, p, length) in enumerate(ref.edges):
            lib().orc_lookup_build(C.byref(mc), n, C.byref(d.c), C.byref(p.c), length,
                                   _up(self.masks), K, _dp(self.lookup[b]))
        return self.lookup

    def preplace(self, seq: str) -> np.ndarray:
        if self.lookup is None:
            self.build_lookup()
        n = self.ref.n
        cols = self.col_tab[np.frombuffer(seq.encode(), dtype=np.uint8)]
        if (cols == 255).any():
            raise ValueError("invalid query character")
        begin, span = valid_range(seq) if self.opts.premasking else (0, n)
        u8 = cols.ctypes.data_as(C.POINTER(C.c_uint8))
        out = np.zeros(len(self.ref.edges))
        for b in range(len(out)):
            out[b] = lib().orc_preplace_score(_dp(self.lookup[b]), self.K, u8, begin, span)
        return out

    def thorough(self, seq: str, edge: int) -> Placement:
        ref = self.ref
        d, p, length = ref.edges[edge]
        m = self.mask_tab[np.frombuffer(seq.encode(), dtype=np.uint8)].astype(np.uint32)
        if (m == 0).any():
            raise ValueError("invalid query character")
        begin, span = valid_range(seq) if self.opts.premasking else (0, ref.n)
        if span == 0:
            raise ValueError("query has no non-gap sites")
        res = OrcBlo()
        fn = lib().orc_place_thorough if self.opts.sliding_blo else lib().orc_place_thorough_raxml
        fn(C.byref(self.pmodel.c()), ref.n, C.byref(d.c), C.byref(p.c), length,
           _up(np.ascontiguousarray(m)), begin, span, C.byref(res))
        pl = Placement(edge, res.logl, 0.0, res.pendant, res.distal)
        pl.rounds, pl.restored = res.rounds, res.restored
        return pl

    def candidates(self, pre: np.ndarray):
        if self.opts.heuristic == 1:
            # until_top_percent (set_manipulators.cpp:82-88)
            keep = int(math.ceil(self.opts.prescoring_threshold * len(pre)))
            order = sorted(range(len(pre)), key=lambda i: (-pre[i], i))
            return order[:keep]
        if self.opts.heuristic == 2:
            # baseball_heuristic (src/core/heuristics.hpp:74-117)
            order = sorted(range(len(pre)), key=lambda i: (-pre[i], i))
            thresh = pre[order[0]] - 3.0
            hits = next((k for k, i in enumerate(order) if pre[i] < thresh), len(order))
            return order[:min(len(order), hits + min(40 - hits, 6))]
        lwr = np.zeros_like(pre)
        lib().orc_lwr(_dp(pre), len(pre), _dp(lwr))
        idx = np.zeros(len(pre), dtype=np.int32)
        k = lib().orc_select_accumulated(_dp(lwr), len(pre), self.opts.prescoring_threshold,
                                         idx.ctypes.data_as(C.POINTER(C.c_int)))
        return [int(i) for i in idx[:k]]

    def place(self, seq: str):
        """One query through preplacement -> candidate selection -> thorough -> LWR -> filter."""
        if self.opts.prescoring:
            cand = self.candidates(self.preplace(seq))
        else:
            cand = list(range(len(self.ref.edges)))
        pls = [self.thorough(seq, b) for b in sorted(cand)]
        logl = np.array([p.logl for p in pls])
        lwr = np.zeros_like(logl)
        lib().orc_lwr(_dp(logl), len(logl), _dp(lwr))
        for p, w in zip(pls, lwr):
            p.lwr = float(w)
        pls.sort(key=lambda p: (-p.lwr, p.edge))
        keep = lib().orc_filter_support(_dp(np.array([p.lwr for p in pls])), len(pls),
                                        self.opts.support_threshold, self.opts.filter_min,
                                        self.opts.filter_max)
        return pls[:keep]


def run_files(tree_file, ref_msa, query_file, model_desc, opts: Options = None, per_rate=None, bugcompat=True):
    """Whole-run restatement of main.cpp:470-540 for unrooted trees; returns
    ({name: [Placement]}, numbered newick, Placer)."""
    opts = opts or Options()
    model = parse_model(model_desc)
    rn, rs = read_fasta(ref_msa)
    qn, qs = read_fasta(query_file)
    if opts.premasking:
        mask = gap_mask(rs) | gap_mask(qs)
        rs, qs = apply_mask(rs, mask), apply_mask(qs, mask)
    tree = build_tree(open(tree_file).read())
    model.per_rate_scalers = (len(tree.tips) > 2000) if per_rate is None else per_rate
    # the reference reads per-rate scalers of a window at a wrong offset: default = as it does
    model.bugcompat_focus = bool(bugcompat) and model.per_rate_scalers
    ref = Reference(tree, model, rn, rs)
    placer = Placer(ref, opts)
    out = {}
    for name, seq in zip(qn, qs):
        out[name] = placer.place(seq)
    return out, numbered_newick(tree), placer


# --------------------------------------------------------------------------------------------
#  jplace helpers
# --------------------------------------------------------------------------------------------
def read_jplace(path: str):
    """{name: [[edge, logl, lwr, distal, pendant], ...]}, tree string."""
    doc = json.load(open(path))
    out = {}
    for pq in doc["placements"]:
        for name in pq["n"]:
            out[name] = pq["p"]
    return out, doc["tree"]


def ref_binary() -> str:
    return os.path.join(HERE, "_ref", "epa-ng")


def run_reference(tree_file, ref_msa, query_file, model_desc, outdir, threads=1, extra=()):
    """Runs the unmodified reference binary built by oracle/Makefile.ref."""
    os.makedirs(outdir, exist_ok=True)
    cmd = [ref_binary(), "-t", tree_file, "-s", ref_msa, "-q", query_file, "-m", model_desc,
           "-w", outdir, "-T", str(threads), "--redo", *extra]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return read_jplace(os.path.join(outdir, "epa_result.jplace"))
