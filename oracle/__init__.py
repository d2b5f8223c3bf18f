"""CPU oracle for the SNAC-24k token->waveform hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``project_morpheus_b200/`` may import
this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and there
only as the checker / the reported CPU baseline, never as the product path.

PARITY UNPINNED.  The arithmetic of the reference path lives in the third-party
PyPI package ``snac`` (``requirements.txt:8`` -> ``snac>=1.2.1,<2``; model id
``hubertsiuzdak/snac_24khz``), which is neither vendored under
``/root/reference`` nor installed in this image, and the reference's own tests
stub it out (``tests/conftest.py:16-30``) and hold no golden vector for
``convert_to_audio`` / ``tokens_decoder``.  The oracle is therefore

(``tests/test_oracle_vs_dac.py`` additionally checks the shared building blocks - Snake1d and the whole
DecoderBlock - against the independent DAC implementation shipped in the ``transformers`` wheel.)

* ``snac_ref``        - a restatement of the published SNAC-24k decode algorithm
                        (quantizer ``from_codes`` + decoder), anchored on the
                        reference's single call site
                        ``Morpheus_Client/tts_engine/speechpipe.py:118``,
* ``speechpipe_ref``  - an independent restatement of the reference's own
                        Python semantics (``speechpipe.py:64-137,146-189,191-337``),
                        proven byte-identical to the verbatim reference file in
                        the authoring container (``tests/test_oracle_vs_reference.py``),
* ``ref_loader``      - runs the verbatim reference ``speechpipe.py`` (when
                        ``/root/reference`` is mounted) with ``snac_ref`` injected as
                        the ``snac`` module, the same trick the reference's
                        ``tests/test_speechpipe_snac_path.py:7-33`` uses.
"""
