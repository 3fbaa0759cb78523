"""CPU oracle for the WorldForge guided-sampling hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU in plain PyTorch,
the algorithm of the reference's per-step denoising loop (Wan2.1 DiT forward,
3D-VAE round trip, UniPC/FLF scheduler step, IRR re-noise, DSG) so the CUDA
path in ``worldforge_b200`` can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; the product path never does.

Pinning: the reference ships no tests or golden vectors.
The oracle is pinned instead against the reference itself, imported from
``/root/reference`` in the build container by ``oracle/make_golden.py``, whose
outputs are committed under ``tests/golden/``.
"""
