"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds CPU restatements of the reference algorithms on the SA-M4C hot path
 plus the loader that imports the UNMODIFIED reference from
/root/reference (available in the build container only) to pin those restatements.

Nothing under `sam_textvqa_b200/` may import this package.  Allowed importers:
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` -- as the checker or the timed CPU baseline, never as the product.

Pinning status: the reference ships no tests or golden vectors, so
parity is pinned by outputs of the reference itself, generated here by
`oracle/make_golden.py` (unmodified `/root/reference/sam/*.py` run through the
`oracle/shim` stand-ins for the two un-vendored third-party packages) and committed as
`tests/golden/*.npz`.
"""
