"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the batched PC-SAFT phase-equilibrium hot path: a restatement of the reference's
Helmholtz models and Newton-step property formulas (``oracle.eos``), a CPU solver standing in for
the un-vendored ``feos`` crate (``oracle.solver``), the public-API-shaped model classes built from
them (``oracle.models``) and a shim that imports the unmodified reference Python from
``/root/reference`` where it exists (``oracle.refshim``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``feos_torch_b200``) never does and has no CPU fallback.
"""
