"""CPU oracle for the fish-diffusion hot path -- TEST INFRASTRUCTURE ONLY.

A plain-numpy restatement of the reference algorithm (fishaudio/fish-diffusion @ 8e8f8cd) for the path
named by BASELINE.json: WaveNet denoiser, DDPM/PLMS/UniPC samplers, NSF-HiFiGAN generator and the mel
front end.  Every function cites the reference file:line it follows.

Rules (task spec, section 3):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this package;
  * the product (``fish_diffusion_b200``) never imports it and has no CPU path at all;
  * pinning: the reference ships NO golden vectors / assertion-bearing tests for this path,
    so the oracle is pinned against outputs of the *reference modules themselves*, imported by file path
    from /root/reference in the build container by ``tests/golden/make_golden.py``; the resulting vectors
    are committed under ``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.
    Pieces the reference cannot execute here (``GaussianDiffusion`` needs mmengine, the mel front end needs
    librosa) are pinned through their importable parts (noise predictors, UniPC, torch.stft) -- see
    DESIGN.md "Oracle pinning" for the exact list and what remains "parity unpinned".

Default arithmetic is float64 (the arbiter); pass ``dtype=np.float32`` where a bit-faithful fp32 comparison with
the reference's own fp32 CPU path is wanted (schedule tables, SineGen increments).
"""
