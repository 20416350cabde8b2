"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the BBBP multi-input network hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the CPU arm being timed.
The product (``bbbp_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Where the arithmetic lives: the reference delegates every op on the path to stock
PyTorch (``torch.nn`` / ``torch.optim.AdamW``; no version pinned by the reference, see
SURVEY.md section 8c).  The oracle is therefore "torch CPU fp32 as installed in this
image" (torch 2.11.0), restated in ``oracle/nets.py`` and pinned two ways:

* against the reference's own classes, AST-extracted from ``/root/reference`` when that
  tree is present (``oracle/reference_classes.py``; the build container only), and
* against golden vectors committed under ``tests/golden/`` that were generated from
  those reference classes by ``oracle/make_golden.py`` (including known-answer vectors
  from the two shipped checkpoints ``best_nn_model*.pth`` and ``maccs_pca.pkl``).

Parity status: MLP family and the PCA projection are pinned by reference artefacts;
the transformer-CNN variants are pinned only by reference-class outputs on seeded
random-init weights (the reference ships no trained transformer-CNN weights).
"""
