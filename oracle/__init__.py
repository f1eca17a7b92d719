"""CPU oracle for the enhancement signal path.  TEST INFRASTRUCTURE ONLY.

This package restates, on plain CPU PyTorch / NumPy, the arithmetic of the
reference's hot path so the CUDA kernels can be checked
against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or the timed CPU baseline -- never as part of the product path.
The product package (``speech_enhancement_by_s3prl_b200``) must not import
anything from here and raises when its CUDA library is missing.

Parity status
-------------
* In-repo arithmetic (``objective.SISDR/L1/WSD``, ``evaluation.sisdr_eval``,
  ``utils.masked_mean/masked_normalize_decibel``, ``model.Linear/LinearResidual``,
  ``dataset.add_noise/normalize_wav_decibel/collate_fn``,
  ``runner._get_length_masks/_decode_wav``): restated in ``oracle/signal_path.py``
  and PINNED against the unmodified reference modules imported in the build
  container through ``oracle/ref_loader.py`` -- the outputs are committed as
  ``tests/golden/*.npz`` by ``oracle/make_golden.py``.
* ``OnlinePreprocessor`` (STFT / iSTFT / features): lives in the un-vendored,
  un-pinned S3PRL dependency (reference README.md:12-27), absent from
  ``/root/reference``.  ``oracle/preprocessor.py`` restates it on
  ``torch.stft`` / ``torch.istft`` from the call-site evidence of SURVEY.md
  Appendix B.  For that component parity is UNPINNED by any reference test or
  fixture (the reference has none); it is anchored instead on an independent
  float64 direct-DFT restatement (``oracle/stft_f64.py``) and on torchaudio's
  ``compute_deltas`` / ``melscale_fbanks``.
"""
