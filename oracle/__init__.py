"""oracle/ — CPU restatement of the reference's algorithm for the hot path (TEST INFRASTRUCTURE, not product).

Plain PyTorch fp32 functional code over a reference-layout state_dict, written from the reference's module
definitions with file:line citations (paths relative to the reference repo root).  It exists because the
reference itself cannot travel to the GPU box.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package consistencytta_b200/ never does.

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is pinned
against outputs of the reference modules themselves, run in the build container by oracle/make_golden.py
(which imports /root/reference/easy_inference) and committed under tests/golden/.
"""
