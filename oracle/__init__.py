"""CPU oracle for the neural-audio-fp inference path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy, with torch-CPU for
the convolutions) of the reference's fingerprinting + retrieval arithmetic.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it -- and there only as the checker / timed CPU
baseline, never as the product.  The product path (``neural-audio-fp_b200``) never
imports it and fails loudly when the CUDA library is missing.

PARITY UNPINNED: the reference's arithmetic lives in un-vendored third-party packages
(tensorflow 2.4.1, kapre 0.3.5, librosa 0.8.1, faiss 1.6.5 -- ``environment.yml``) that are
not installable here (no network), and the reference tree has no golden fingerprints,
spectrograms or hit-rate tables.  The restatement follows the reference call sites file
by file (cited in each function) and the published algorithms of those packages; it is
cross-checked against independent implementations available in this image
(torch.stft, torchaudio.melscale_fbanks, torch.nn.functional.{conv2d,layer_norm},
scipy.spatial.distance.cdist) and against the few known answers the reference does hold
(parameter counts ``model/fp/nnfp.py:270-274``, segment counts
``model/utils/audio_utils.py:173-177``, ``eval/test_ids_icassp2021.npy``).
"""
