"""Minimal stand-ins for the gymnasium types the reference's env surface uses (gymnasium is not installed
in this image).  When the real package is importable it is used instead, so SB3 sees genuine spaces."""
try:  # pragma: no cover - exercised only where gymnasium exists
    from gymnasium import spaces  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # ModuleNotFoundError in this image
    from . import spaces  # noqa: F401
    HAVE_GYMNASIUM = False
