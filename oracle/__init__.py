"""CPU oracle for the LI-VAE rVAE/VAE training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as the thing shipped.  The product path
(``li-vae_b200/``) never imports this package and fails loudly when its CUDA
extension is missing.

What it is: a functional restatement (plain torch-CPU / numpy, no nn.Module,
no autograd.Function) of the reference algorithm for the hot path named in
SURVEY.md section 8, each function citing the reference file:line it follows.
The reference's arithmetic lives in PyTorch ATen ops (torch 2.9.1 pinned in the
reference's uv.lock; this image has torch 2.11.0), so the restatement calls the
same dense ATen ops (conv2d, linear, max_pool2d, interpolate, pad) and restates
by hand only the pieces the CUDA kernels re-derive: affine_grid + grid_sample
(bilinear / reflection / align_corners=False) forward and backward, the
reparameterisation, the ELBO reductions and the patch crops.

Pinning: the reference's own tests hold NO golden vector for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the reference
itself, imported unmodified from /root/reference in the build container by
``tests/golden/make_golden.py`` (committed, with the vectors it wrote under
``tests/golden/``).  ``tests/test_oracle_golden.py`` checks every oracle
function against those vectors on CPU.
"""
