"""CPU oracle for the MSI inference hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference's algorithm
(brownvc/matryodshka @ 831c407: geometry/{spherical,sampling,projector}.py,
matryodshka/{msi,nets}.py), written op-for-op in float32 NumPy (geometry) and
torch-CPU float32 (conv net).  Every function cites the reference file:line it
follows.

It is NOT part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
only as the checker / the reported CPU baseline.  The product path
(``matryodshka_b200``) never imports it and fails loudly when its CUDA library
is missing.

PINNED TO THE REFERENCE'S OWN CODE, OVER A RESTATED OP LAYER.  The reference ships no tests, golden
vectors or fixtures and TensorFlow 1.14 cannot be installed here, but its Python
source runs under Python 3: oracle/refrun/run_reference.py imports the UNMODIFIED files
/root/reference/geometry/{spherical,projector,sampling}.py and matryodshka/{msi,nets}.py over a NumPy
stand-in for the TensorFlow ops they call (oracle/refrun/tf114_numpy.py), drives them as test.py:110-170
does and writes tests/golden/reference_run.npz.  tests/test_reference_golden.py holds this restatement to
those vectors: geometry (sweep coordinates, validity mask, PSV, all four renderers, depth, uint8 output,
jittered sweep, resampling) BIT FOR BIT, including SHA-256 digests at the full 320x640x32 size; the conv
nets (coord and wrap-pad, all four colour schemes, ngf 8 and 64) within 1e-5.  So the reference's own
algorithm -- signs, swaps, bracketing, index / channel / argument order, layer wiring, scope names -- is
pinned by execution, not by reading.

What remains restated (and is marked [TF-1.14] at the point of use, here and in tf114_numpy.py) is the
arithmetic of the third-party dependency that is absent from /root/reference: tensorflow==1.14.0
(LinSpace, matmul summation order, SAME padding, slim.conv2d / conv2d_transpose / layer_norm, gather_nd,
floor-mod, convert_image_dtype) and tensorflow-graphics==1.0.0 (rotation_matrix_3d.from_euler).  No
TensorFlow binary is available to check those against; beyond the reference run the oracle is held by
analytic known-answer tests (tests/test_oracle_kat.py) and an independent float64 twin.
"""

from . import geometry_np, net_torch, msi_np  # noqa: F401
