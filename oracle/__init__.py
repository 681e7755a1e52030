"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32/fp64) restatement of the reference algorithm on the hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product (nextgen_uia_b200) never does.  Parity pinning: the reference ships no golden vectors or
tests, so the oracle is pinned against the reference's OWN modules executed in the build
container (oracle/make_golden.py imports /root/reference/src/adapters/{mona,lora}.py and
src/losses/losses.py by path and writes tests/golden/*.pt); the timm / open_clip / HF-BERT arithmetic is
restated from the pinned versions' published semantics ("parity unpinned" for those third-party parts:
no source under /root/reference, nothing installed to run — see DESIGN.md).
"""
