"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's classification-by-ELBO hot path, used solely as the checker:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything from here.  The product package (diffusion-classifier_b200/dcb200) never does.

Parity status
  * loop (classify / schedule / q_sample / prune): PINNED against the reference's verbatim code
    (oracle/reference_loader.py imports /root/reference/diffusion/diffusion_classifier.py with stub
    modules; oracle/make_golden.py writes tests/golden/*.npz from it; tests/test_oracle_*.py compare).
  * denoisers (diffusers 0.31.0 U-Net / DiT) and Haar (pywt): PARITY UNPINNED -- third-party code that
    is neither vendored in /root/reference nor installable offline; restated from the published
    algorithm and anchored on closed-form known-answer tests.
"""
