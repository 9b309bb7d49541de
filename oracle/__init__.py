"""CPU oracle for the Simple-RF per-ray rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``simple_rf_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` do, and there only as the checker / the timed CPU baseline.

What it is: a functional (no ``nn.Module``) restatement in CPU torch ops of the reference
algorithm, stage by stage, each function citing the reference ``file:line`` it follows
(paths relative to the upstream checkout).  The reference is 100 % PyTorch, so the
restatement uses the same ATen CPU kernels for ``sum`` / ``cumsum`` / ``searchsorted`` /
``grid_sample``; ``oracle/aten_order.py`` additionally spells the exact fp32/fp64
evaluation order of those kernels out in numpy, because that order is what the CUDA
kernels reproduce for the bit-exact contracts (sample indices, occupancy mask, compaction).

Parity pinning: the reference ships no tests and no golden vectors.
The oracle is therefore pinned against *outputs of the reference itself*:
``oracle/generate_golden.py`` imports the unmodified reference modules from the upstream
checkout (only possible in the build container) with fixed seeds, checks every oracle stage
against them, and writes the small fixtures committed under ``tests/golden/``.
"""
