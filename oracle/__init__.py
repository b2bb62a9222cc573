"""CPU oracle for the VAEL logic hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under oracle/ restates, on the CPU, what the reference (EleMisi/VAEL) computes
on this path, each function citing the reference file:line it follows.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it,
and only as the checker or the timed CPU baseline; vael_b200/ never does.

Pinning: the restatement is checked against golden vectors produced by importing the
reference itself in the build container (tests/golden/make_golden.py, committed with the
.npz fixtures it wrote).  What is pinned that way: GraphSemiring ops incl. the batch-wide
`.any()` shortcuts, the dense closed form (worlds, query, gradients), facts head, Herbrand
base, w_q, gumbel-softmax(hard) with injected noise, the ELBO terms, the Mario matrix path
and the four Mario tables.  What is NOT pinned: problog's `sdd.evaluate` traversal itself
(problog / PySDD are unpinned third-party packages absent from this image) — "parity
unpinned" for that traversal; oracle/graph_path.py follows its published semantics
 over this repo's own compiled circuit.
"""
