"""B200-native SNAC-24k token->waveform engine behind Morpheus's ``speechpipe`` interface."""
__version__ = "0.1.0"
