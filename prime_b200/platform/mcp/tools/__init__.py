from . import availability, pods, ssh  # noqa: F401
