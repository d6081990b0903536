"""Which reference source lines the critical path of a compiled circuit runs through (build container only: traces the
unmodified reference from /root/reference).  usage: critical_path.py CASE [prefix]   (CASE: a key of tests/golden/make_golden.py CASES)"""
import collections
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg  # noqa: E402
from bounty_matrix_inversion_b200.fhe import tracing  # noqa: E402
from bounty_matrix_inversion_b200.fhe.program import lower  # noqa: E402

name = sys.argv[1]
fn, inputset, _ins, _meta = mg.CASES[name]()
sites = {}
orig = tracing.Trace.new_lookup


def patched(self, src, f, group, vals):
    r = orig(self, src, f, group, vals)
    ref = [fr for fr in traceback.extract_stack(limit=40) if "/root/reference" in fr.filename]
    sites[self.jobs[-1].base] = " < ".join(f"{os.path.basename(fr.filename)}:{fr.lineno}" for fr in reversed(ref[-3:]))
    return r


tracing.Trace.new_lookup = patched
enc = {"arrays": "encrypted", "signs": "encrypted"} if name.startswith("qf") else {"x": "encrypted", "y": "encrypted"}
trace, outs, _shapes, _ = mg.fhe.Compiler(fn, enc).trace(inputset)
prog = lower(trace, outs, (len(outs),), collapse_borrows="prefix" if "prefix" in sys.argv[2:] else True)
level_of, job_info = prog.debug["level_of"], prog.debug["job_info"]
site_of = {}
for jb in trace.jobs:                      # representative -> site of the first traced lookup it stands for
    sub = prog.debug["subst"].get(jb.base)
    rep = jb.base if sub is None else (next(iter(sub)) if len(sub) == 1 else None)
    if rep is not None:
        site_of.setdefault(rep, sites[jb.base])
for helper, base in prog.debug["split_of"].items():
    site_of.setdefault(helper, "split: " + sites[base])
end = max(level_of, key=level_of.get)
path, b = [], end
while b is not None:
    path.append(b)
    preds = [t for t, _c in job_info[b][0][0] if t in level_of]
    b = max(preds, key=level_of.get) if preds else None
print(f"{name}: {prog.n_pbs} lookups, {len(prog.levels)} levels; critical path by reference call site:")
for site, cnt in collections.Counter(site_of.get(b, "(derived: prefix-scan sign digit)") for b in path).most_common(12):
    print(f"{cnt:6d}  {site}")

