"""Host-side data formats either side of the hot path: FLAC decode and the
Kaldi-compatible log-mel front-end used by ``extract_feature.py`` / the s3prl expert."""
from .flac import decode_flac  # noqa: F401
