"""CPU oracle for the CS-ViT hot path - TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker or the timed CPU baseline.  Nothing under ``cs-vit_b200/``
imports it; the product path has no CPU fallback.

What it restates (plain fp32 PyTorch on CPU, each function citing the lines it follows):

* ``swin_restated``  - the Swin-v1 forward of HuggingFace ``transformers`` (pinned 4.51.3 by the reference's
  ``poetry.lock``, 5.5.0 installed here), the third-party code the reference's backbone call
  ``self.backbone(imgs_norm).last_hidden_state`` (ref:cs_vit/net/ti_poser.py:246,426) actually executes.
* ``head_restated``  - the CS-ViT head: perspective encoder, spatial encoder, temporal encoders, output heads,
  6D->axis-angle, forward kinematics and loss (ref:cs_vit/net/ti_poser.py, ref:cs_vit/net/transformer_module.py,
  ref:cs_vit/utils/geometry.py, ref:cs_vit/utils/joint.py).

Pinning.  The reference ships no tests, golden vectors or known answers for this path:
"parity unpinned" by the reference itself.  The oracle is instead pinned against OUTPUTS OF THE REFERENCE RUN
HERE: ``make_goldens.py`` imports the real ``cs_vit.net.ti_poser.Poser`` from ``/root/reference`` (with stubs for
the packages this image lacks) and HF ``SwinModel``, runs them on the seeded synthetic inputs, checks the
restatement against them to fp32 round-off, and commits the reference's outputs under ``tests/golden/``.
``tests/test_oracle.py`` re-checks restatement-vs-golden on every run without needing ``/root/reference``.
The MANO layer is a synthetic stand-in on both sides (the real one is licence-gated, see
``cs_vit/utils/mano_standin.py``).
"""
