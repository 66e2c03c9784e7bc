"""CPU oracle for the SVDD decoding hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch fp32 / numpy) restatement of the reference's
algorithm for the path behind ``decode.py`` / ``decode_tweedie.py``
(masa-ue/SVDD).  Every function cites the reference file:line it follows.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``svdd_b200/`` imports it; the product path has no CPU
    fallback and raises if the CUDA library is missing.

Pinning status
--------------
  * stages 2 and 4 (SUBS, q_xs, Gumbel-max draw, selection), the noise
    schedule, the CNN denoiser, the ConvGRU value net and the full SVDD-MC /
    SVDD-PM loops are PINNED: ``tests/golden/make_golden.py`` imports the
    reference's own ``diffusion_gosai.py`` / ``models/dnaconv.py`` /
    ``noise_schedule.py`` / ``Enformer.py`` from ``/root/reference`` (through
    import stubs for the absent third-party packages), runs them on seeded
    inputs and commits the outputs under ``tests/golden/``; the oracle is
    checked bit-exactly (integer outputs) / to 1e-6 (fp32 outputs) against them
    in ``tests/test_oracle_golden.py``.
  * the Enformer-style DNA value net is "parity unpinned" at ONE boundary: its
    ``Attention`` / ``AttentionPool`` / ``GELU`` / ``relative_shift`` /
    ``exponential_linspace_int`` live in the un-vendored, un-pinned third-party
    package ``enformer_pytorch`` (absent from /root/reference and from this
    image).  ``oracle/enformer_shim.py`` restates the published algorithm
    (lucidrains/enformer-pytorch ``modeling_enformer.py``; cross-checked against
    the commented restatement in the reference at ``Enformer.py:2659-2768``).
    Everything *around* those five symbols (``EnformerTrunk``, ``ConvBlock``,
    ``ConvHead`` ...) is pinned by running the reference's own ``Enformer.py``
    on top of the shim.
"""
