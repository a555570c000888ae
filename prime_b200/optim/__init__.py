from .schedule import lr_at  # noqa: F401
