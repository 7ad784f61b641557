from .expert import UpstreamExpert  # noqa: F401
