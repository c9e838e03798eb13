"""Diagnostic (CPU, oracle): smoothing rounds and Newton evaluations per thorough pair on a
cfg2-like synthetic dataset. Test infrastructure only."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle(); orc.build()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ds = pkg.synth.dataset(T=T, n_sites=1000, n_queries=NQ, window=200)
tree = orc.build_tree(ds["newick"])
model = orc.parse_model(ds["model"])
seqs = [bytes(r).decode() for r in ds["ref"]]
ref = orc.Reference(tree, model, ds["names"], seqs)
pl = orc.Placer(ref)
t0 = time.time(); pl.build_lookup(); print("lookup", time.time() - t0)
cnt = C.c_ulonglong.in_dll(orc.lib(), "orc_stat_deriv_calls")
rounds, evals, restored, ncand = [], [], 0, []
for q in ds["queries"]:
    seq = bytes(q).decode()
    cand = pl.candidates(pl.preplace(seq))
    ncand.append(len(cand))
    for e in cand:
        cnt.value = 0
        p = pl.thorough(seq, e)
        rounds.append(p.rounds); evals.append(cnt.value); restored += p.restored
rounds, evals = np.array(rounds), np.array(evals)
print("pairs", len(rounds), "cand/query", np.mean(ncand))
print("rounds mean %.2f hist %s" % (rounds.mean(), np.bincount(rounds)))
print("deriv evals/pair mean %.1f median %.0f p90 %.0f max %d; per round %.1f" % (evals.mean(), np.median(evals), np.percentile(evals, 90), evals.max(), evals.sum() / rounds.sum()))
print("restored", restored)
H = (C.c_ulonglong * 40).in_dll(orc.lib(), "orc_stat_nr_hist")
print("nr calls", C.c_ulonglong.in_dll(orc.lib(), "orc_stat_nr_calls").value, "clamped steps", C.c_ulonglong.in_dll(orc.lib(), "orc_stat_clamped").value)
print("evals reaching iteration k:", list(H)[:34])
