"""CPU oracle for the Tiny-NewsRec hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch-fp32 / numpy restatement of the reference's
algorithm for the path named in BASELINE.json (news encoder -> gathers -> user
encoder -> click scoring + CE -> multi-teacher KD loss -> Adam(amsgrad); news
table build; impression scoring with AUC/MRR/nDCG).  Every function cites the
reference file:line it follows (paths relative to /root/reference).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it, and only as the checker / CPU baseline.
The product package (`tiny-newsrec_b200/`) never imports it and has no CPU
fallback: it raises if the CUDA extension is missing.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself,
imported in the build container through `tests/golden/ref_shim.py`; the
generated vectors are committed under `tests/golden/` together with the script
(`tests/golden/make_golden.py`) and checked by `tests/test_oracle_golden.py`.
"""
from . import model, metrics, batching, optim  # noqa: F401
